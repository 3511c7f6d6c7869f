// bf16 tcgen05 GEMM building blocks of the COPE_PREC_BF16 MLP path (declarations; kernels in tc_gemm.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace cope {

typedef __nv_bfloat16 bf16;

enum TcEpi : int {
  TC_STORE = 0,        // v = alpha*(acc + bias?)                          n<nsplit -> out ; else out2[n-nsplit]
  TC_BIAS_SOFTPLUS,    // out = alpha*softplus100(acc+bias)
  TC_BIAS_RELU,        // out = max(acc+bias,0)
  TC_BIAS_SIGMOID,     // out = sigmoid(acc+bias)
  TC_MUL_SIGP,         // n<nsplit: out = alpha*acc*sp(H) ; else out2[n-nsplit] = alpha*acc
  TC_TANGENT,          // sp = sp(H); out = alpha*acc*sp ; out2 = acc*D*100*(1-sp)
  TC_BWD,              // a' = acc + r1[m]*r1w[n]; n<nsplit: out = alpha*a'*sp(H) + D? ; else out2[n-nsplit] = alpha*a'
  TC_RELU_MASK,        // out = (H>0) ? acc + r1.. : 0
};
// sp(H) = softplus'(z) recovered from the stored activation h = softplus(z):  1 - exp(-100 * H * hscale)

struct TcArgs {
  int M, N, K;                 // rows; padded N (mult of 16, <= 256); padded K (mult of 64, <= 320)
  const bf16* A; int lda;      // activations [M x lda] bf16, K-contiguous
  const bf16* Bp;              // packed weights [K/8][N][8]
  int epi; float alpha; float hscale;
  int n_valid;                 // columns >= n_valid are not stored
  int nsplit;                  // see TcEpi (>= N: no split)
  const float* bias;           // [>= n_valid] or null
  void* out; int ldo; int out_f32;
  void* out2; int ldo2; int out2_f32; int n2_valid;
  const bf16* H; int ldh;
  const bf16* D; int ldd;
  const float* r1; int r1_ld; const float* r1w;   // rank-1 update acc += r1[m*r1_ld] * r1w[n]
  int stages;                  // A-ring depth (set by launch_tc_gemm)
};
inline TcArgs tc_args(int M, int N, int K, const bf16* A, int lda, const bf16* Bp, void* out, int ldo, int out_f32) {
  TcArgs t{};
  t.M = M; t.N = N; t.K = K; t.A = A; t.lda = lda; t.Bp = Bp; t.out = out; t.ldo = ldo; t.out_f32 = out_f32;
  t.epi = TC_STORE; t.alpha = 1.0f; t.hscale = 1.0f; t.n_valid = N; t.nsplit = 1 << 30; t.n2_valid = 1 << 30;
  return t;
}
int launch_tc_gemm(const TcArgs& a, cudaStream_t s);

// dW[m, n] += sum_p X[p, m] * Y[p, n]   for up to two (X, Y) operand pairs; m < m_valid, n < n_valid
struct TcWgradArgs {
  int64_t P;
  int Mp, Np;                  // padded widths: Mp in {128, 256}, Np mult of 16 <= 256
  int m_valid, n_valid;
  const bf16* X[2]; int ldx[2];
  const bf16* Y[2]; int ldy[2];
  int n_pairs;
  float* dW; int ldw;
  float* part;                 // workspace: tc_wgrad_part_floats() floats (per-CTA partial tiles + column sums)
  float* db;                   // optional: db[m] += sum_p X[0][p, m]  (bias gradient, fused column sum of pair 0's X)
  int atomic;                  // set by launch_tc_wgrad: CTAs add their tile straight into dW / db with L2 reductions
};
int64_t tc_wgrad_part_floats();
int launch_tc_wgrad(const TcWgradArgs& a, cudaStream_t s);
int launch_tc_wgrad_batch(const TcWgradArgs* probs, int n, cudaStream_t s);

// packed[(k/8) * Np + n][k%8] = W[sn * ldw + sk] (transposed: W[sk * ldw + sn]) where sn / sk are the source indices
// that the segment lists map packed row n / packed k to (unmapped -> 0).  Jobs are queued into a PackBatch and
// flushed as ONE kernel launch (all layers of a network).
struct PackSeg { int dst, src, len; };
struct PackSpec {
  PackSeg n[2]; int nn;
  PackSeg k[4]; int nk;
};
inline PackSpec pack_spec() { PackSpec p{}; return p; }
inline PackSpec& seg_n(PackSpec& p, int dst, int src, int len) { p.n[p.nn++] = PackSeg{dst, src, len}; return p; }
inline PackSpec& seg_k(PackSpec& p, int dst, int src, int len) { p.k[p.nk++] = PackSeg{dst, src, len}; return p; }
constexpr int kMaxPackJobs = 24;
struct PackJob { const float* W; bf16* out; PackSpec spec; int ldw, Np, Kp, transposed, first; };
struct PackBatch { PackJob jobs[kMaxPackJobs]; int n, total; };
inline void pack_add(PackBatch& b, const float* W, int ldw, const PackSpec& spec, int Np, int Kp, int transposed, bf16* out) {
  PackJob& j = b.jobs[b.n++];
  j.W = W; j.out = out; j.spec = spec; j.ldw = ldw; j.Np = Np; j.Kp = Kp; j.transposed = transposed; j.first = b.total;
  b.total += Np * Kp;
}
int launch_tc_pack_batch(const PackBatch& b, cudaStream_t s);
int launch_tc_pack(const float* W, int ldw, const PackSpec& spec, int Np, int Kp, int transposed, bf16* out, cudaStream_t s);

}  // namespace cope
