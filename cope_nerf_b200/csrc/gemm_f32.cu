// fp32 SIMT GEMM with fused NeuS epilogues — the strict-parity (COPE_PREC_FP32) MLP path.
// C[M x N] = epi( op(A) [M x K] * op(B) [K x N] ).  128x128x16 tiles, 256 threads, 8x8 per thread,
// register-prefetched global loads, split-K with atomics for the weight-gradient (K = points) shape.
#include <stdarg.h>
#include <algorithm>

#include "common.cuh"

namespace cope {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error_str() { return g_err; }

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int PAD = 4;

template <bool TR>  // TR=false: src is [rows x K] (k contiguous); TR=true: src is [K x rows] (row idx contiguous)
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int ld, int row0, int nrows, int k0,
                                          int kend, float (&r)[8]) {
  const int t = threadIdx.x;
  if (!TR) {
    const int m = row0 + (t >> 1);
    const int kb = k0 + (t & 1) * 8;
    const float* p = src + (int64_t)m * ld + kb;
    const bool mok = m < nrows;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (mok && kb + j < kend) ? __ldg(p + j) : 0.0f;
  } else {
    const int k = k0 + (t >> 4);
    const int mb = row0 + (t & 15) * 8;
    const float* p = src + (int64_t)k * ld + mb;
    const bool kok = k < kend;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (kok && mb + j < nrows) ? __ldg(p + j) : 0.0f;
  }
}
template <bool TR>
__device__ __forceinline__ void store_tile(float (*sm)[BM + PAD], const float (&r)[8]) {
  const int t = threadIdx.x;
  if (!TR) {
    const int m = t >> 1, kb = (t & 1) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[kb + j][m] = r[j];
  } else {
    const int k = t >> 4, mb = (t & 15) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[k][mb + j] = r[j];
  }
}

__device__ __forceinline__ void epilogue(const GemmArgs& a, int m, int n, float acc) {
  const int64_t ci = (int64_t)m * a.ldc + n;
  switch (a.epi) {
    case EPI_STORE: {
      float v = a.alpha * acc + (a.bias ? a.bias[n] : 0.0f);
      if (n < a.nsplit) a.C[ci] = v;
      else a.C2[(int64_t)m * a.ldc2 + (n - a.nsplit)] = v;
    } break;
    case EPI_ATOMIC: atomicAdd(a.C + ci, a.alpha * acc); break;
    case EPI_BIAS_SOFTPLUS: {
      float z = acc + a.bias[n];
      if (a.C2) a.C2[(int64_t)m * a.ldc2 + n] = z;
      a.C[ci] = a.alpha * act_fwd(a.act, a.act_slope, z);
    } break;
    case EPI_BIAS_RELU: a.C[ci] = fmaxf(acc + a.bias[n], 0.0f); break;
    case EPI_BIAS_SIGMOID: a.C[ci] = sigmoidf_(acc + a.bias[n]); break;
    case EPI_MUL_SIGP:
      if (n < a.nsplit) a.C[ci] = a.alpha * acc * act_d1(a.act, a.act_slope, a.Z[(int64_t)m * a.ldz + n]);
      else a.C2[(int64_t)m * a.ldc2 + (n - a.nsplit)] = a.alpha * acc;
      break;
    case EPI_TANGENT: {
      float sp = softplus100_d1(a.Z[(int64_t)m * a.ldz + n]);
      a.C[ci] = a.alpha * acc * sp;
      a.C2[(int64_t)m * a.ldc2 + n] = acc * a.D[(int64_t)m * a.ldd + n] * (kSoftplusBeta * (1.0f - sp));
    } break;
    case EPI_BWD:
      if (n < a.nsplit) {
        float v = a.alpha * acc * act_d1(a.act, a.act_slope, a.Z[(int64_t)m * a.ldz + n]);
        if (a.D) v += a.D[(int64_t)m * a.ldd + n];
        a.C[ci] = v;
      } else if (a.C2) {
        a.C2[(int64_t)m * a.ldc2 + (n - a.nsplit)] = a.alpha * acc;
      }
      break;
    case EPI_RELU_MASK: a.C[ci] = a.Z[(int64_t)m * a.ldz + n] > 0.0f ? acc : 0.0f; break;
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(NT) sgemm_kernel(const GemmArgs a) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  int kbeg = 0, kend = a.K;
  if (a.split_k > 1) {
    int chunk = (int)(((int64_t)a.K + a.split_k - 1) / a.split_k);
    chunk = (chunk + BK - 1) / BK * BK;
    kbeg = blockIdx.z * chunk;
    kend = min(a.K, kbeg + chunk);
    if (kbeg >= kend) return;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[8];
  // A: TA=false -> [M x K]; B: TB=true -> [N x K] (same access pattern as non-transposed A)
  load_tile<TA>(a.A, a.lda, m0, a.M, kbeg, kend, ra);
  load_tile<!TB>(a.B, a.ldb, n0, a.N, kbeg, kend, rb);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    store_tile<TA>(As, ra);
    store_tile<!TB>(Bs, rb);
    __syncthreads();
    if (k0 + BK < kend) {
      load_tile<TA>(a.A, a.lda, m0, a.M, k0 + BK, kend, ra);
      load_tile<!TB>(a.B, a.ldb, n0, a.N, k0 + BK, kend, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < a.N) epilogue(a, m, n, acc[i][j]);
    }
  }
}

// Small-problem variant: 32 x 64 x 16 tiles, 256 threads, 2 x 4 outputs per thread.  The MotionNetwork runs on ~1000 time samples
// (M ~ 990, N = K = 256): the 128 x 128 kernel covers that with 16 CTAs on 148 SMs and ~60 us of exposed latency per GEMM; this one
// launches 124.  Same operand conventions and epilogues.
constexpr int SM_ = 32, SN_ = 64, SK_ = 16;
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_small_kernel(const GemmArgs a) {
  __shared__ float As[SK_][SM_ + 1];
  __shared__ float Bs[SK_][SN_ + 1];
  const int m0 = blockIdx.y * SM_, n0 = blockIdx.x * SN_;
  int kbeg = 0, kend = a.K;
  if (a.split_k > 1) {
    int chunk = (int)(((int64_t)a.K + a.split_k - 1) / a.split_k);
    chunk = (chunk + SK_ - 1) / SK_ * SK_;
    kbeg = blockIdx.z * chunk;
    kend = min(a.K, kbeg + chunk);
    if (kbeg >= kend) return;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
  for (int k0 = kbeg; k0 < kend; k0 += SK_) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {                       // A tile: 32 x 16
      const int idx = threadIdx.x + 256 * i;
      const int mm = TA ? (idx % SM_) : (idx / SK_), kk = TA ? (idx / SM_) : (idx % SK_);
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.0f;
      if (m < a.M && k < kend) v = TA ? a.A[(int64_t)k * a.lda + m] : a.A[(int64_t)m * a.lda + k];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // B tile: 16 x 64
      const int idx = threadIdx.x + 256 * i;
      const int nn = TB ? (idx / SK_) : (idx % SN_), kk = TB ? (idx % SK_) : (idx / SN_);
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.0f;
      if (n < a.N && k < kend) v = TB ? a.B[(int64_t)n * a.ldb + k] : a.B[(int64_t)k * a.ldb + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK_; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = Bs[k][tx * 4 + j];
        acc[0][j] = fmaf(a0, b, acc[0][j]);
        acc[1][j] = fmaf(a1, b, acc[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < a.N) epilogue(a, m, n, acc[i][j]);
    }
  }
}

int launch_gemm(bool transA, bool transB, const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0) return 0;
  COPE_REQUIRE(a.K > 0, "gemm: K must be positive (M=%d N=%d K=%d)", a.M, a.N, a.K);
  COPE_REQUIRE(a.split_k == 1 || a.epi == EPI_ATOMIC, "gemm: split-K needs the atomic epilogue");
  const int64_t big_ctas = (int64_t)((a.N + BN - 1) / BN) * ((a.M + BM - 1) / BM) * a.split_k;
  if (big_ctas < 74) {               // less than half a wave of the 128 x 128 tiles: the small-tile kernel fills the machine
    dim3 grid((a.N + SN_ - 1) / SN_, (a.M + SM_ - 1) / SM_, a.split_k);
    if (!transA && transB) sgemm_small_kernel<false, true><<<grid, 256, 0, s>>>(a);
    else if (!transA && !transB) sgemm_small_kernel<false, false><<<grid, 256, 0, s>>>(a);
    else if (transA && !transB) sgemm_small_kernel<true, false><<<grid, 256, 0, s>>>(a);
    else sgemm_small_kernel<true, true><<<grid, 256, 0, s>>>(a);
    COPE_CHECK_LAUNCH("sgemm_small");
    return 0;
  }
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, a.split_k);
  if (!transA && transB) sgemm_kernel<false, true><<<grid, NT, 0, s>>>(a);
  else if (!transA && !transB) sgemm_kernel<false, false><<<grid, NT, 0, s>>>(a);
  else if (transA && !transB) sgemm_kernel<true, false><<<grid, NT, 0, s>>>(a);
  else sgemm_kernel<true, true><<<grid, NT, 0, s>>>(a);
  COPE_CHECK_LAUNCH("sgemm");
  return 0;
}

}  // namespace cope

extern "C" {
int cope_version(void) { return 100; }
const char* cope_last_error(void) { return cope::last_error_str(); }
uint64_t cope_launch_count(void) { return cope::g_launches; }

int cope_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, int accumulate, cope_stream_t s) {
  using namespace cope;
  GemmArgs g = gemm_args(M, N, K, A, lda, B, ldb, C, ldc);
  if (accumulate) {
    g.epi = EPI_ATOMIC;
    int tiles = (int)(ceil_div(M, BM) * ceil_div(N, BN));
    g.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(K, 4 * BK), ceil_div(2 * 148, tiles)));
  }
  return launch_gemm(transA != 0, transB != 0, g, as_stream(s));
}
}
