// Shared helpers for libcope_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cope_b200.h"

namespace cope {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)

#define COPE_CHECK_LAUNCH(name)                                              \
  do {                                                                       \
    ++cope::g_launches;                                                      \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      cope::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return -2;                                                             \
    }                                                                        \
  } while (0)

#define COPE_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      cope::set_error(__VA_ARGS__);    \
      return -1;                       \
    }                                  \
  } while (0)

static inline cudaStream_t as_stream(cope_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kSoftplusBeta = 100.0f;

// softplus(beta=100) with torch's threshold (beta*x > 20 -> x)           model/neus_fields.py:266
__device__ __forceinline__ float softplus100(float z) {
  float bz = kSoftplusBeta * z;
  return bz > 20.0f ? z : log1pf(expf(bz)) * (1.0f / kSoftplusBeta);
}
// d softplus / dz = sigmoid(100 z) (1 above the threshold, as torch's softplus_backward)
__device__ __forceinline__ float softplus100_d1(float z) {
  float bz = kSoftplusBeta * z;
  if (bz > 20.0f) return 1.0f;
  float e = expf(bz);
  return e / (e + 1.0f);
}
// hidden activation selected by cope_mlp_desc::activation (softplus(beta=100) or LeakyReLU(slope))
__device__ __forceinline__ float act_fwd(int act, float slope, float z) {
  return act == COPE_ACT_LEAKY_RELU ? (z > 0.0f ? z : slope * z) : softplus100(z);
}
__device__ __forceinline__ float act_d1(int act, float slope, float z) {
  return act == COPE_ACT_LEAKY_RELU ? (z > 0.0f ? 1.0f : slope) : softplus100_d1(z);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- fp32 GEMM with fused epilogues (gemm_f32.cu) -------------------------------------------------
enum GemmEpi : int {
  EPI_STORE = 0,        // v = alpha*acc (+bias[n]); n<nsplit: C[m,n] = v; else C2[m,n-nsplit] = v
  EPI_ATOMIC,           // C += alpha*acc (atomicAdd; used with split-K)
  EPI_BIAS_SOFTPLUS,    // z = acc+bias; Z=z; C = alpha*softplus100(z)
  EPI_BIAS_RELU,        // C = max(acc+bias, 0)
  EPI_BIAS_SIGMOID,     // C = sigmoid(acc+bias)
  EPI_MUL_SIGP,         // n<nsplit: C = alpha*acc*softplus'(Z[m,n]);  n>=nsplit: C2[m,n-nsplit] = alpha*acc
  EPI_TANGENT,          // u=acc; sp=softplus'(Z); C = alpha*u*sp; C2 = u*D*100*(1-sp)
  EPI_BWD,              // n<nsplit: C = alpha*acc*softplus'(Z) + D[m,n] (D may be null); n>=nsplit: C2 = alpha*acc
  EPI_RELU_MASK,        // C = acc * (Z[m,n] > 0)
};

struct GemmArgs {
  int M, N, K;
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  int epi;
  float alpha;
  const float* bias;
  const float* Z; int ldz;     // aux input 1
  const float* D; int ldd;     // aux input 2
  float* C2; int ldc2;         // aux output
  int nsplit;                  // column split for skip-layer epilogues (>= N: no split)
  int split_k;                 // >1 only with EPI_ATOMIC
  int act; float act_slope;    // activation of the *_SOFTPLUS / *_SIGP / BWD epilogues (COPE_ACT_*; default softplus(beta=100))
};

int launch_gemm(bool transA, bool transB, const GemmArgs& a, cudaStream_t s);
inline GemmArgs gemm_args(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.epi = EPI_STORE; g.alpha = 1.0f; g.nsplit = 1 << 30; g.split_k = 1;
  return g;
}

}  // namespace cope
