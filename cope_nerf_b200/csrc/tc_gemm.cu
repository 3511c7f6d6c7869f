// bf16 tcgen05 GEMMs for the NeuS MLPs (sm_100a).
//
// tc_gemm_kernel<EPI>  C[M x N] = epi(A[M x K] * W[N x K]^T): persistent CTAs, one 128-row tile at a time.
//   * W (<= 160 KB, pre-packed in UMMA core-matrix order) is bulk-TMA'd into shared memory ONCE per CTA and
//     stays resident; A streams through a 4-stage cp.async ring (128 x 64 bf16 per stage);
//   * one elected thread issues tcgen05.mma (M=128, N<=256, K=16) into one of two TMEM accumulators, so the
//     epilogue of tile i overlaps the MMAs of tile i+1;
//   * 8 epilogue warps pull 32-column slabs with tcgen05.ld, prefetch the next slab's auxiliary operands from
//     global memory, and apply the NeuS epilogue selected at compile time (bias / rank-1 vectors live in smem).
// tc_wgrad_kernel  dW[Mp x Np] += X^T Y with K = points (both operands MN-major): each CTA accumulates its slice of
//   points in TMEM (fp32) and writes ONE partial tile; wgrad_reduce_kernel sums the partials (deterministic, no
//   atomics).
//
// Shared-memory operand layout: no swizzle, 8 x 16 B core matrices (see tc_common.cuh::smem_desc).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "tc_common.cuh"
#include "tc_gemm.cuh"
#include "sdf_fused.cuh"

namespace cope {
using namespace tc;

constexpr int kTcThreads = 320;          // warps 0-7 epilogue, warp 8 TMA producer, warp 9 MMA issuer + TMEM owner
constexpr int kMmaWarp = 9;
constexpr int kStages = 8;               // maximum ring depth; the launch picks as many as fit next to W
constexpr int kTileM = 128;
constexpr int kChunkK = 64;
constexpr int kAStageBytes = kTileM * kChunkK * 2;   // 16 KB: TMA box 64 x 128, 128B-swizzled
constexpr int kVecBytes = 2 * 256 * 4;               // bias + rank-1 row staged in smem

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus(beta=100): log2(1 + 2^(100 log2e z)) * ln2/100, torch's threshold (100 z > 20 -> z); 2 MUFU + 4 FP ops
__device__ __forceinline__ float fast_softplus100(float z) {
  const float t = z * (kSoftplusBeta * 1.4426950408889634f);
  const float s = lg2f(1.0f + ex2f(t)) * (0.6931471805599453f / kSoftplusBeta);
  return t > 28.853900817779268f ? z : s;
}
__device__ __forceinline__ float sp_from_h(float h, float hscale) {   // softplus'(z) = 1 - exp(-100 softplus(z))
  return 1.0f - ex2f(h * (hscale * -kSoftplusBeta * 1.4426950408889634f));
}
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = bf16_lo(q.x); f[1] = bf16_hi(q.x); f[2] = bf16_lo(q.y); f[3] = bf16_hi(q.y);
  f[4] = bf16_lo(q.z); f[5] = bf16_hi(q.z); f[6] = bf16_lo(q.w); f[7] = bf16_hi(q.w);
}
struct Aux32 { uint4 q[4]; };   // 32 bf16
__device__ __forceinline__ void load_aux32(const bf16* p, Aux32& a) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) a.q[i] = __ldg(q + i);
}

// full 32-wide store (n0 multiple of 32, row pointer 16 B aligned)
__device__ __forceinline__ void store32(void* base, int64_t off, int is_f32, const float (&v)[32]) {
  if (is_f32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      q.x = pack_bf16(v[8 * i], v[8 * i + 1]); q.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
      q.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]); q.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
      o[i] = q;
    }
  }
}
__device__ __noinline__ void store_elem(void* base, int64_t off, int is_f32, float v) {
  if (is_f32) reinterpret_cast<float*>(base)[off] = v;
  else reinterpret_cast<bf16*>(base)[off] = __float2bfloat16(v);
}

// One 32-column slab [n0, n0+32) of row m.  h/d: auxiliary operands already in registers.
template <int EPI>
__device__ __forceinline__ void tc_epilogue32(const TcArgs& a, const float* s_bias, const float* s_r1w, int64_t m, int n0,
                                              float (&acc)[32], const Aux32& h, const Aux32& d, float r1) {
  constexpr bool kNeedH = EPI == TC_MUL_SIGP || EPI == TC_TANGENT || EPI == TC_BWD || EPI == TC_RELU_MASK;
  constexpr bool kNeedD = EPI == TC_TANGENT || EPI == TC_BWD;
  constexpr bool kSplit = EPI == TC_STORE || EPI == TC_MUL_SIGP || EPI == TC_BWD;
  float o2[EPI == TC_TANGENT ? 32 : 1];
  float hv[kNeedH ? 32 : 1], dv[kNeedD ? 32 : 1];
  if (kNeedH) {
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack8(h.q[i], hv + 8 * i);
  }
  const bool haveD = kNeedD && (EPI == TC_TANGENT || a.D != nullptr);
  if (kNeedD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack8(d.q[i], dv + 8 * i);
  }
  // `lo` = this slab lies entirely below nsplit (the common case): no per-element split test
  const bool lo = !kSplit || n0 + 32 <= a.nsplit;
  const bool use_r1 = EPI == TC_BWD && a.r1 != nullptr;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int n = n0 + i;
    float v = acc[i];
    if (EPI <= TC_BIAS_SIGMOID) v += s_bias[n];
    if (EPI == TC_BWD) { if (use_r1) v += r1 * s_r1w[n]; }
    if (EPI == TC_STORE) v *= a.alpha;
    else if (EPI == TC_BIAS_SOFTPLUS) v = a.alpha * fast_softplus100(v);
    else if (EPI == TC_BIAS_RELU) v = fmaxf(v, 0.0f);
    else if (EPI == TC_BIAS_SIGMOID) v = __fdividef(1.0f, 1.0f + ex2f(v * -1.4426950408889634f));
    else if (EPI == TC_MUL_SIGP) {
      const float s = a.alpha * v;
      v = (lo || n < a.nsplit) ? s * sp_from_h(hv[kNeedH ? i : 0], a.hscale) : s;
    } else if (EPI == TC_TANGENT) {
      const float sp = sp_from_h(hv[kNeedH ? i : 0], a.hscale);
      o2[EPI == TC_TANGENT ? i : 0] = v * dv[kNeedD ? i : 0] * (kSoftplusBeta * (1.0f - sp));
      v = a.alpha * v * sp;
    } else if (EPI == TC_BWD) {
      const float s = a.alpha * v;
      if (lo || n < a.nsplit) {
        v = s * sp_from_h(hv[kNeedH ? i : 0], a.hscale);
        if (haveD) v += dv[kNeedD ? i : 0];
      } else {
        v = s;
      }
    } else if (EPI == TC_RELU_MASK) v = hv[kNeedH ? i : 0] > 0.0f ? v : 0.0f;
    acc[i] = v;
  }
  if constexpr (EPI == TC_TANGENT) {
    if (n0 + 32 <= a.n_valid) {
      store32(a.out, m * a.ldo + n0, a.out_f32, acc);
      store32(a.out2, m * a.ldo2 + n0, a.out2_f32, o2);
    } else {
      for (int i = 0; i < 32; ++i)
        if (n0 + i < a.n_valid) {
          store_elem(a.out, m * a.ldo + n0 + i, a.out_f32, acc[i]);
          store_elem(a.out2, m * a.ldo2 + n0 + i, a.out2_f32, o2[i]);
        }
    }
    return;
  } else {
  const int lim1 = kSplit ? min(a.n_valid, a.nsplit) : a.n_valid;     // columns < lim1 -> out
  if (n0 + 32 <= lim1 && ((a.ldo * (a.out_f32 ? 4 : 2)) % 16 == 0)) {
    store32(a.out, m * a.ldo + n0, a.out_f32, acc);
  } else if (kSplit && a.out2 && n0 >= a.nsplit && (n0 - a.nsplit) + 32 <= a.n2_valid &&
             (((n0 - a.nsplit) * (a.out2_f32 ? 4 : 2)) % 16 == 0) && ((a.ldo2 * (a.out2_f32 ? 4 : 2)) % 16 == 0)) {
    store32(a.out2, m * a.ldo2 + (n0 - a.nsplit), a.out2_f32, acc);
  } else {
    for (int i = 0; i < 32; ++i) {
      const int n = n0 + i;
      if (n < lim1) store_elem(a.out, m * a.ldo + n, a.out_f32, acc[i]);
      else if (kSplit && a.out2 && n >= a.nsplit && (n - a.nsplit) < a.n2_valid)
        store_elem(a.out2, m * a.ldo2 + (n - a.nsplit), a.out2_f32, acc[i]);
    }
  }
  }
}

template <int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmA) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr bool kNeedH = EPI == TC_MUL_SIGP || EPI == TC_TANGENT || EPI == TC_BWD || EPI == TC_RELU_MASK;
  constexpr bool kNeedD = EPI == TC_TANGENT || EPI == TC_BWD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbytes = (uint32_t)a.K * a.N * 2;
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((wbytes + 1023) & ~1023u);
  const int nstages = a.stages;
  float* s_bias = reinterpret_cast<float*>(sA + nstages * kAStageBytes);
  float* s_r1w = s_bias + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_r1w + 256);
  uint64_t* w_full = bars;                 // [5] one per 64-wide K slab of W
  uint64_t* a_full = bars + 5;
  uint64_t* a_empty = bars + 5 + kStages;
  uint64_t* acc_full = bars + 5 + 2 * kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (threadIdx.x == 0) {
    for (int k = 0; k < 5; ++k) mbar_init(w_full + k, 1);
    for (int s = 0; s < nstages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, 8); }
    fence_barrier_init();
  }
  if (threadIdx.x < 256) {
    s_bias[threadIdx.x] = (a.bias && (int)threadIdx.x < a.n_valid) ? a.bias[threadIdx.x] : 0.0f;
    s_r1w[threadIdx.x] = (a.r1w && (int)threadIdx.x < a.N) ? a.r1w[threadIdx.x] : 0.0f;
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntiles = (a.M + kTileM - 1) / kTileM;
  const int nkc = a.K / kChunkK;

  if (warp == 8) {
    // ------------------------------------------------------------------ producer: weights once (bulk TMA), A ring
    // (2-D tiled TMA, 64 x 128 bf16 boxes, 128B swizzle; rows past M are zero-filled by the tensor map)
    if (threadIdx.x == 256) {
      tma_prefetch_desc(&tmA);
      const uint32_t slab = (uint32_t)a.N * 128;          // one K chunk of W: [8 k-cores][N][8] bf16
      for (int k = 0; k < nkc; ++k) {
        mbar_arrive_expect_tx(w_full + k, slab);
        bulk_g2s(sW + k * slab, reinterpret_cast<const uint8_t*>(a.Bp) + (size_t)k * slab, slab, w_full + k);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int kc = 0; kc < nkc; ++kc) {
          mbar_wait(a_empty + stage, phase ^ 1);
          mbar_arrive_expect_tx(a_full + stage, kAStageBytes);
          tma_load_2d(sA + stage * kAStageBytes, &tmA, kc * kChunkK, tile * kTileM, a_full + stage);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(kTileM, a.N, 0, 0);
      bool w_ready = false;
      const uint32_t sWa = smem_u32(sW);
      const uint32_t b_lbo = (uint32_t)a.N * 16;     // K-adjacent cores of W: one full [N][8] slab apart
      int stage = 0;
      uint32_t phase = 0, accp = 0;
      int acc = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(acc_empty + acc, accp ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kc = 0; kc < nkc; ++kc) {
          if (!w_ready) mbar_wait(w_full + kc, 0);          // first tile only: this K slab of W has landed
          mbar_wait(a_full + stage, phase);
          tc_fence_after();
          const uint32_t sAa = smem_u32(sA + stage * kAStageBytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks) {
            const uint64_t ad = smem_desc_sw128(sAa + ks * 32, 16, 1024);
            const uint64_t bd = smem_desc(sWa + (uint32_t)(kc * 8 + ks * 2) * b_lbo, b_lbo, 128);
            umma_bf16(d_tmem, ad, bd, idesc, (kc | ks) != 0);
          }
          umma_commit(a_empty + stage);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + acc);
        w_ready = true;
        acc ^= 1;
        if (acc == 0) accp ^= 1;
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps 0..7
    const int q = warp & 3, half = warp >> 2;
    const int nch = (a.N + 31) / 32;
    uint32_t accp = 0;
    int acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t m = (int64_t)tile * kTileM + q * 32 + lane;
      const bool row_ok = m < a.M;
      const int64_t mm = row_ok ? m : 0;
      Aux32 h{}, d{};
      int c = half;
      // aux operands of the first slab: in flight while the MMAs of this tile finish
      if (c < nch) {
        if (kNeedH && c * 32 < a.nsplit) load_aux32(a.H + mm * a.ldh + c * 32, h);
        if (kNeedD && a.D && c * 32 < a.nsplit) load_aux32(a.D + mm * a.ldd + c * 32, d);
      }
      const float r1 = (EPI == TC_BWD && a.r1) ? a.r1[mm * a.r1_ld] : 0.0f;
      mbar_wait(acc_full + acc, accp);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      for (; c < nch; c += 2) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        Aux32 hn{}, dn{};
        const int cn = c + 2;
        if (cn < nch) {    // prefetch the next slab's aux operands
          if (kNeedH && cn * 32 < a.nsplit) load_aux32(a.H + mm * a.ldh + cn * 32, hn);
          if (kNeedD && a.D && cn * 32 < a.nsplit) load_aux32(a.D + mm * a.ldd + cn * 32, dn);
        }
        if (row_ok) tc_epilogue32<EPI>(a, s_bias, s_r1w, m, c * 32, v, h, d, r1);
        h = hn; d = dn;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
      acc ^= 1;
      if (acc == 0) accp ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

static int tc_gemm_stages(int N, int K) {
  size_t w = ((size_t)K * N * 2 + 1023) & ~(size_t)1023;
  const int fit = (int)((232448 - (int64_t)w - kVecBytes - 256) / kAStageBytes);
  return std::max(1, std::min(kStages, fit));
}
static size_t tc_gemm_smem(int N, int K) {
  size_t w = ((size_t)K * N * 2 + 1023) & ~(size_t)1023;
  return w + tc_gemm_stages(N, K) * kAStageBytes + kVecBytes + 256;
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time libcuda dependency) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
struct MapKey {
  const void* p; uint64_t cols, rows, ld; uint32_t box_rows;
  bool operator==(const MapKey& o) const { return p == o.p && cols == o.cols && rows == o.rows && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.p);
    h ^= k.cols * 0x9E3779B97F4A7C15ull + k.rows * 0xC2B2AE3D27D4EB4Full + k.ld * 0x165667B19E3779F9ull + k.box_rows;
    return h;
  }
};
// bf16 matrix [rows x cols] (row stride ld elements), boxes of 64 columns x box_rows rows, 128B swizzle, OOB -> 0
static int make_tmap(const bf16* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  MapKey key{ptr, cols, rows, ld, box_rows};
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return 0; }
  EncodeTiledFn fn = encode_fn();
  COPE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  COPE_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, "TMA operand must be 16-byte aligned (ld=%llu)",
               (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  COPE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] ld %llu", (int)r, (unsigned long long)rows,
               (unsigned long long)cols, (unsigned long long)ld);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// bf16 [layers][rows][ld] buffer (cols addressed), boxes of 64 columns x 128 rows x 1 layer, 128B swizzle; rows past
// `rows` are zero-filled on load and clipped on store, so a partial last tile never touches the next layer
int make_tmap3(const bf16* ptr, uint64_t cols, uint64_t rows, uint64_t layers, uint64_t ld, uint64_t layer_stride, CUtensorMap* out) {
  struct Key3 { const void* p; uint64_t c, r, l, ld, ls; };
  static std::vector<std::pair<Key3, CUtensorMap>> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  for (auto& e : cache)
    if (e.first.p == ptr && e.first.c == cols && e.first.r == rows && e.first.l == layers && e.first.ld == ld && e.first.ls == layer_stride) {
      *out = e.second;
      return 0;
    }
  EncodeTiledFn fn = encode_fn();
  COPE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  COPE_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0 && (layer_stride * 2) % 16 == 0,
               "TMA operand must be 16-byte aligned (ld=%llu)", (unsigned long long)ld);
  cuuint64_t dims[3] = {cols, rows, layers};
  cuuint64_t strides[2] = {ld * 2, layer_stride * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  COPE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed (%d) for [%llu x %llu x %llu] ld %llu", (int)r,
               (unsigned long long)layers, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  if (cache.size() > 256) cache.clear();
  cache.emplace_back(Key3{ptr, cols, rows, layers, ld, layer_stride}, *out);
  return 0;
}

template <int EPI>
static int launch_tc_gemm_t(const TcArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    COPE_REQUIRE(e == cudaSuccess, "tc_gemm: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int ntiles = (a.M + kTileM - 1) / kTileM;
  TcArgs b = a;
  b.stages = tc_gemm_stages(a.N, a.K);
  CUtensorMap tmA;
  if (int rc = make_tmap(a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, kTileM, &tmA)) return rc;
  tc_gemm_kernel<EPI><<<std::min(ntiles, 148), kTcThreads, tc_gemm_smem(a.N, a.K), s>>>(b, tmA);
  COPE_CHECK_LAUNCH("tc_gemm");
  return 0;
}

int launch_tc_gemm(const TcArgs& a, cudaStream_t s) {
  if (a.M <= 0) return 0;
  COPE_REQUIRE(a.N % 16 == 0 && a.N >= 16 && a.N <= 256, "tc_gemm: N=%d must be a multiple of 16 in [16,256]", a.N);
  COPE_REQUIRE(a.K % kChunkK == 0 && a.K >= kChunkK && a.K <= 320, "tc_gemm: K=%d must be a multiple of 64 in [64,320]", a.K);
  COPE_REQUIRE(a.lda % 8 == 0 && a.lda >= a.K, "tc_gemm: lda=%d must be >= K and a multiple of 8", a.lda);
  COPE_REQUIRE(tc_gemm_smem(a.N, a.K) <= 232448, "tc_gemm: N=%d K=%d does not fit shared memory", a.N, a.K);
  const int n32 = ((a.N + 31) / 32) * 32;
  const bool needH = a.epi == TC_MUL_SIGP || a.epi == TC_TANGENT || a.epi == TC_BWD || a.epi == TC_RELU_MASK;
  const int aux_w = a.nsplit >= n32 ? n32 : std::min(n32, ((a.nsplit + 31) / 32) * 32);   // columns the epilogue reads
  COPE_REQUIRE(!needH || (a.H && a.ldh % 8 == 0 && a.ldh >= aux_w), "tc_gemm: aux H missing / too narrow (ldh=%d)", a.ldh);
  COPE_REQUIRE(!a.D || (a.ldd % 8 == 0 && a.ldd >= aux_w), "tc_gemm: aux D too narrow (ldd=%d)", a.ldd);
  switch (a.epi) {
    case TC_STORE: return launch_tc_gemm_t<TC_STORE>(a, s);
    case TC_BIAS_SOFTPLUS: return launch_tc_gemm_t<TC_BIAS_SOFTPLUS>(a, s);
    case TC_BIAS_RELU: return launch_tc_gemm_t<TC_BIAS_RELU>(a, s);
    case TC_BIAS_SIGMOID: return launch_tc_gemm_t<TC_BIAS_SIGMOID>(a, s);
    case TC_MUL_SIGP: return launch_tc_gemm_t<TC_MUL_SIGP>(a, s);
    case TC_TANGENT: return launch_tc_gemm_t<TC_TANGENT>(a, s);
    case TC_BWD: return launch_tc_gemm_t<TC_BWD>(a, s);
    case TC_RELU_MASK: return launch_tc_gemm_t<TC_RELU_MASK>(a, s);
  }
  COPE_REQUIRE(false, "tc_gemm: unknown epilogue %d", a.epi);
}

// ================================================================================================ wgrad
constexpr int kWgStages = 3;
constexpr int kWgChunkP = 64;            // points per pipeline stage
constexpr int kWgBoxBytes = 64 * kWgChunkP * 2;   // one TMA box: 64 columns x 64 points, 128B-swizzled (8 KB)

struct WgMaps { CUtensorMap x[2], y[2]; };
// several weight-gradient problems in ONE launch: problem i owns CTAs [cta0[i], cta0[i+1]); far fewer (and fatter) point
// slices per problem than one launch each, so the per-CTA tile flush is paid ~18 instead of 148 times per layer
constexpr int kWgMaxProb = 12;
struct WgBatch { TcWgradArgs p[kWgMaxProb]; int cta0[kWgMaxProb + 1]; int n; };
struct WgBatchMaps { WgMaps m[kWgMaxProb]; };
// workspace = 148 partial tiles [256 x 256] + 148 partial column-sum rows [256]
__host__ __device__ constexpr int64_t tc_wgrad_part_floats_c() { return (int64_t)148 * 256 * 256 + (int64_t)148 * 256; }
int64_t tc_wgrad_part_floats() { return tc_wgrad_part_floats_c(); }

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgBatch batch,
                                                                 const __grid_constant__ WgBatchMaps bmaps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int pi = 0;
  while (pi + 1 < batch.n && (int)blockIdx.x >= batch.cta0[pi + 1]) ++pi;
  const TcWgradArgs& a = batch.p[pi];
  const WgMaps& maps = bmaps.m[pi];
  const int slice = (int)blockIdx.x - batch.cta0[pi], nslices = batch.cta0[pi + 1] - batch.cta0[pi];
  // stage = X tile (Mp/64 boxes) followed by Y tile (Np/64 boxes); box b of a tile holds columns [64b, 64b+64)
  const int xb = a.Mp >> 6, yb = a.Np >> 6;
  const uint32_t xbytes = xb * kWgBoxBytes, ybytes = yb * kWgBoxBytes;
  const uint32_t stage_bytes = xbytes + ybytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  uint64_t* s_full = bars;
  uint64_t* s_empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(s_full + s, 1); mbar_init(s_empty + s, 1 + 8); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's slice of points (multiples of the stage size); the host sizes the grid so every CTA has work
  const int64_t nchunks_total = (a.P + kWgChunkP - 1) / kWgChunkP;
  const int64_t per = (nchunks_total + nslices - 1) / nslices;
  const int64_t ch0 = (int64_t)slice * per, ch1 = min(nchunks_total, ch0 + per);
  const int nmb = a.Mp / 128;
  const bool have_work = ch0 < ch1;

  if (threadIdx.x == 256) {
    // ------------------------------------------------------------------ TMA producer (points past P are zero-filled)
    int stage = 0;
    uint32_t phase = 0;
    for (int pr = 0; pr < a.n_pairs; ++pr) {
      tma_prefetch_desc(&maps.x[pr]);
      tma_prefetch_desc(&maps.y[pr]);
      for (int64_t ch = ch0; ch < ch1; ++ch) {
        mbar_wait(s_empty + stage, phase ^ 1);
        mbar_arrive_expect_tx(s_full + stage, stage_bytes);
        uint8_t* sx = smem + stage * stage_bytes;
        uint8_t* sy = sx + xbytes;
        const int p0 = (int)(ch * kWgChunkP);
        for (int b = 0; b < xb; ++b) tma_load_2d(sx + b * kWgBoxBytes, &maps.x[pr], b * 64, p0, s_full + stage);
        for (int b = 0; b < yb; ++b) tma_load_2d(sy + b * kWgBoxBytes, &maps.y[pr], b * 64, p0, s_full + stage);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0 && have_work) {
      const uint32_t idesc = idesc_bf16(128, a.Np, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int pr = 0; pr < a.n_pairs; ++pr) {
        for (int64_t ch = ch0; ch < ch1; ++ch) {
          mbar_wait(s_full + stage, phase);
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + stage * stage_bytes), sy = sx + xbytes;
#pragma unroll
          for (int ks = 0; ks < kWgChunkP / 16; ++ks) {
            // MN-major, 128B swizzle: 64-column blocks kWgBoxBytes apart, 8-point groups 1024 B apart, 16 points per MMA
            const uint64_t bd = smem_desc_sw128(sy + ks * 2048, kWgBoxBytes, 1024);
            for (int mb = 0; mb < nmb; ++mb) {
              const uint64_t ad = smem_desc_sw128(sx + ks * 2048 + mb * 2 * kWgBoxBytes, kWgBoxBytes, 1024);
              umma_bf16(tmem_base + mb * 256, ad, bd, idesc, first ? 0u : 1u);
            }
            first = false;
          }
          umma_commit(s_empty + stage);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit(acc_full);
    }
  } else if (warp < 8 && have_work) {
    // ---- while the MMAs run: bias gradient = column sums of pair 0's X tile, read straight from the swizzled smem
    // stage (thread t owns column t); every warp also releases the stage (s_empty counts MMA commit + 8 warps)
    {
      const int col = threadIdx.x;                       // 0..255
      float csum = 0.0f;
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t cofs = (uint32_t)(col >> 6) * kWgBoxBytes + (uint32_t)(col & 7) * 2;
      const int cchunk = (col & 63) >> 3;
      for (int pr = 0; pr < a.n_pairs; ++pr) {
        for (int64_t ch = ch0; ch < ch1; ++ch) {
          mbar_wait(s_full + stage, phase);
          if (pr == 0 && a.db != nullptr && col < a.Mp) {
            const uint8_t* sx = smem + stage * stage_bytes + cofs;
#pragma unroll 8
            for (int p = 0; p < kWgChunkP; ++p)
              csum += __bfloat162float(*reinterpret_cast<const bf16*>(sx + p * 128 + ((cchunk ^ (p & 7)) << 4)));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty + stage);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
      if (a.db != nullptr && col < a.Mp) {
        if (a.atomic) { if (col < a.m_valid) atomicAdd(a.db + col, csum); }
        else a.part[tc_wgrad_part_floats_c() - (int64_t)148 * 256 + (size_t)slice * 256 + col] = csum;
      }
    }
    // ---- partial tile of this CTA -> part[blockIdx.x][Mp][Np] (plain stores; wgrad_reduce_kernel sums them)
    const int q = warp & 3, half = warp >> 2;
    const int nch = a.Np / 16;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* part = a.part + (size_t)slice * a.Mp * a.Np;
    const bool vec_ok = (a.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0;
    for (int mb = 0; mb < nmb; ++mb) {
      const int m = mb * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + mb * 256 + ((uint32_t)(q * 32) << 16);
      for (int c = half; c < nch; c += 2) {
        float v[16];
        tmem_ld16(taddr + c * 16, v);
        if (a.atomic) {
          // L2 reductions straight into dW: 148 CTAs x one tile each, no partial round trip through HBM
          if (m < a.m_valid) {
            float* o = a.dW + (size_t)m * a.ldw + c * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = c * 16 + 4 * i;
              if (vec_ok && n + 4 <= a.n_valid) {
                atomicAdd(reinterpret_cast<float4*>(o + 4 * i), make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (n + k < a.n_valid) atomicAdd(o + 4 * i + k, v[4 * i + k]);
              }
            }
          }
        } else {
          float4* o = reinterpret_cast<float4*>(part + (size_t)m * a.Np + c * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// dW[m, n] += sum_c part[c][m][n] ;  db[m] += sum_c colpart[c][m]
// One thread owns 4 consecutive n of one row (float4 loads) and keeps 8 partial tiles in flight; the tail block sums
// the bias-gradient column partials.  Np % 4 == 0; n_valid is handled per element.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int nparts, int Mp, int Np, int m_valid, int n_valid,
                                    float* __restrict__ dW, int ldw, float* __restrict__ db) {
  const int nq = (n_valid + 3) >> 2;                       // float4 groups per row
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = m_valid * nq * 4;
  const int nw_pad = (nw + 31) & ~31;                      // whole warps: the shuffles below never meet the bias tail
  if (i >= nw_pad) {
    const int m = i - nw_pad;
    if (db == nullptr || m >= m_valid) return;
    const float* p = part + (tc_wgrad_part_floats_c() - (int64_t)148 * 256) + m;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int c = 0;
    for (; c + 4 <= nparts; c += 4) {
      s0 += p[(size_t)c * 256]; s1 += p[(size_t)(c + 1) * 256]; s2 += p[(size_t)(c + 2) * 256]; s3 += p[(size_t)(c + 3) * 256];
    }
    for (; c < nparts; ++c) s0 += p[(size_t)c * 256];
    db[m] += (s0 + s1) + (s2 + s3);
    return;
  }
  // 4 lanes share one float4 of the output: lane k sums partial tiles k, k+4, ... (8 loads in flight each)
  const bool live = i < nw;
  const int g = live ? i >> 2 : 0, sub = i & 3;
  const int m = g / nq, n = (g - m * nq) * 4;
  if (!live) nparts = 0;
  const float4* p = reinterpret_cast<const float4*>(part + (size_t)m * Np + n);
  const size_t stride = (size_t)Mp * Np / 4;
  float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  int c = sub;
  for (; c + 28 < nparts; c += 32) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + (size_t)(c + 4 * u) * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; c < nparts; c += 4) {
    const float4 v = __ldcs(p + (size_t)c * stride);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (live && sub == 0) {
    float* o = dW + (size_t)m * ldw + n;
    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (n + k < n_valid) o[k] += r[k];
  }
}

static int wgrad_check(const TcWgradArgs& a) {
  COPE_REQUIRE((a.Mp == 128 || a.Mp == 256) && a.Np % 64 == 0 && a.Np >= 64 && a.Np <= 256,
               "tc_wgrad: Mp=%d Np=%d unsupported (Mp in {128,256}, Np multiple of 64)", a.Mp, a.Np);
  for (int i = 0; i < a.n_pairs; ++i)
    COPE_REQUIRE(a.ldx[i] % 8 == 0 && a.ldy[i] % 8 == 0 && a.ldx[i] >= a.Mp && a.ldy[i] >= a.Np,
                 "tc_wgrad: operand %d leading dims (%d,%d) must cover the padded tile (%d,%d)", i, a.ldx[i], a.ldy[i], a.Mp, a.Np);
  return 0;
}
static int wgrad_maps(const TcWgradArgs& a, WgMaps* maps) {
  for (int i = 0; i < 2; ++i) {
    const int q = i < a.n_pairs ? i : 0;
    if (int rc = make_tmap(a.X[q], (uint64_t)a.ldx[q], (uint64_t)a.P, (uint64_t)a.ldx[q], kWgChunkP, &maps->x[i])) return rc;
    if (int rc = make_tmap(a.Y[q], (uint64_t)a.ldy[q], (uint64_t)a.P, (uint64_t)a.ldy[q], kWgChunkP, &maps->y[i])) return rc;
  }
  return 0;
}
static int wgrad_attr() {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    COPE_REQUIRE(e == cudaSuccess, "tc_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  return 0;
}
static size_t wgrad_smem(const TcWgradArgs& a) { return (size_t)kWgStages * ((a.Mp >> 6) + (a.Np >> 6)) * kWgBoxBytes + 256; }

int launch_tc_wgrad(const TcWgradArgs& a, cudaStream_t s) {
  if (a.P <= 0 || a.n_pairs <= 0) return 0;
  if (int rc = wgrad_check(a)) return rc;
  COPE_REQUIRE(a.part != nullptr, "tc_wgrad: partial-sum workspace missing");
  if (int rc = wgrad_attr()) return rc;
  WgBatch* batch = new WgBatch();
  WgBatchMaps* maps = new WgBatchMaps();
  int rc = wgrad_maps(a, &maps->m[0]);
  const int64_t nchunks = (a.P + kWgChunkP - 1) / kWgChunkP;
  int grid = (int)std::min<int64_t>(148, std::max<int64_t>(1, nchunks / 2));
  const int64_t per = (nchunks + grid - 1) / grid;
  grid = (int)((nchunks + per - 1) / per);                   // every CTA owns >= 1 chunk
  // default: L2 reductions into dW (fp32 add order varies run to run); COPE_WGRAD_DETERMINISTIC=1 keeps the per-CTA
  // partial tiles + fixed-order reduction kernel
  // (rows of dW that are not 16-byte aligned would fall back to scalar reductions, which are slower than the reduce kernel)
  batch->p[0] = a;
  batch->p[0].atomic = getenv("COPE_WGRAD_DETERMINISTIC") == nullptr && (a.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0;
  batch->cta0[0] = 0; batch->cta0[1] = grid; batch->n = 1;
  const bool atomic = batch->p[0].atomic != 0;
  if (rc == 0) {
    tc_wgrad_kernel<<<grid, kTcThreads, wgrad_smem(a), s>>>(*batch, *maps);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc_wgrad: launch failed: %s", cudaGetErrorString(e)); rc = -2; }
  }
  delete batch;
  delete maps;
  if (rc) return rc;
  if (atomic) return 0;
  const int n = ((a.m_valid * ((a.n_valid + 3) / 4) * 4 + 31) & ~31) + (a.db ? a.m_valid : 0);
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, s>>>(a.part, grid, a.Mp, a.Np, a.m_valid, a.n_valid, a.dW, a.ldw, a.db);
  COPE_CHECK_LAUNCH("wgrad_reduce");
  return 0;
}

// All problems in one launch (L2 reductions into dW / db); CTAs are shared out in proportion to the bytes each problem
// streams.  COPE_WGRAD_DETERMINISTIC=1 falls back to one launch per problem with the fixed-order reduction.
int launch_tc_wgrad_batch(const TcWgradArgs* probs, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  if (getenv("COPE_WGRAD_DETERMINISTIC") != nullptr || n > kWgMaxProb) {
    for (int i = 0; i < n; ++i)
      if (int rc = launch_tc_wgrad(probs[i], s)) return rc;
    return 0;
  }
  if (int rc = wgrad_attr()) return rc;
  WgBatch* batch = new WgBatch();
  WgBatchMaps* maps = new WgBatchMaps();
  int rc = 0, m = 0;
  double w[kWgMaxProb], wsum = 0.0;
  size_t smem = 0;
  for (int i = 0; i < n && rc == 0; ++i) {
    const TcWgradArgs& a = probs[i];
    if (a.P <= 0 || a.n_pairs <= 0) continue;
    rc = wgrad_check(a);
    if (rc == 0) rc = wgrad_maps(a, &maps->m[m]);
    batch->p[m] = a;
    batch->p[m].atomic = 1;
    w[m] = (double)a.P * (a.Mp + a.Np) * a.n_pairs + 2.0e6;      // streamed elements + a fixed cost for the tile flush
    wsum += w[m];
    smem = std::max(smem, wgrad_smem(a));
    ++m;
  }
  if (rc == 0 && m > 0) {
    int total = 0, nct[kWgMaxProb];
    for (int i = 0; i < m; ++i) {
      const int64_t nchunks = (batch->p[i].P + kWgChunkP - 1) / kWgChunkP;
      nct[i] = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, (int64_t)(148.0 * w[i] / wsum)));
      total += nct[i];
    }
    for (int i = 0; total < 148 && i < 8 * m; ++i) {           // hand the left-over SMs to the heaviest problems
      const int k = i % m;
      const int64_t nchunks = (batch->p[k].P + kWgChunkP - 1) / kWgChunkP;
      if (w[k] * m >= wsum && nct[k] < nchunks) { ++nct[k]; ++total; }
    }
    batch->cta0[0] = 0;
    for (int i = 0; i < m; ++i) batch->cta0[i + 1] = batch->cta0[i] + nct[i];
    batch->n = m;
    tc_wgrad_kernel<<<batch->cta0[m], kTcThreads, smem, s>>>(*batch, *maps);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc_wgrad_batch: launch failed: %s", cudaGetErrorString(e)); rc = -2; }
  }
  delete batch;
  delete maps;
  return rc;
}

// ================================================================================================ packing
__device__ __forceinline__ int map_seg(const PackSeg* s, int n, int i) {
  for (int q = 0; q < n; ++q)
    if (i >= s[q].dst && i < s[q].dst + s[q].len) return s[q].src + (i - s[q].dst);
  return -1;
}
__global__ void tc_pack_kernel(const __grid_constant__ PackBatch b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.total) return;
  int q = 0;
  while (q + 1 < b.n && i >= b.jobs[q + 1].first) ++q;
  const PackJob& j = b.jobs[q];
  const int e = i - j.first;
  const int kk = e & 7, n = (e >> 3) % j.Np, kg = e / (8 * j.Np);
  const int sn = map_seg(j.spec.n, j.spec.nn, n), sk = map_seg(j.spec.k, j.spec.nk, kg * 8 + kk);
  float v = 0.0f;
  if (sn >= 0 && sk >= 0) v = j.transposed ? j.W[(int64_t)sk * j.ldw + sn] : j.W[(int64_t)sn * j.ldw + sk];
  j.out[e] = __float2bfloat16(v);
}

int launch_tc_pack_batch(const PackBatch& b, cudaStream_t s) {
  if (b.n == 0) return 0;
  tc_pack_kernel<<<(b.total + 255) / 256, 256, 0, s>>>(b);
  COPE_CHECK_LAUNCH("tc_pack");
  return 0;
}
int launch_tc_pack(const float* W, int ldw, const PackSpec& spec, int Np, int Kp, int transposed, bf16* out, cudaStream_t s) {
  PackBatch b{};
  pack_add(b, W, ldw, spec, Np, Kp, transposed, out);
  return launch_tc_pack_batch(b, s);
}

}  // namespace cope

// ---- raw entry points used by the kernel-level tests ------------------------------------------------
using namespace cope;
extern "C" {

int cope_tc_pack(const float* W, int ldw, int n_src, int k_src, int Np, int Kp, int transposed, void* out_bf16,
                 cope_stream_t s) {
  PackSpec sp = pack_spec();
  seg_n(sp, 0, 0, n_src);
  seg_k(sp, 0, 0, k_src);
  return launch_tc_pack(W, ldw, sp, Np, Kp, transposed, reinterpret_cast<bf16*>(out_bf16), as_stream(s));
}

int cope_tc_gemm(int M, int N, int K, const void* A_bf16, int lda, const void* Bp_bf16, const float* bias, int epi,
                 float alpha, void* out, int ldo, int out_f32, cope_stream_t s) {
  TcArgs t = tc_args(M, N, K, reinterpret_cast<const bf16*>(A_bf16), lda, reinterpret_cast<const bf16*>(Bp_bf16), out, ldo,
                     out_f32);
  t.bias = bias; t.epi = epi; t.alpha = alpha;
  return launch_tc_gemm(t, as_stream(s));
}

int64_t cope_tc_wgrad_ws_floats(void) { return tc_wgrad_part_floats(); }

int cope_tc_wgrad(int64_t P, int Mp, int Np, int m_valid, int n_valid, const void* X, int ldx, const void* Y, int ldy,
                  float* dW, int ldw, float* ws, cope_stream_t s) {
  TcWgradArgs w{};
  w.P = P; w.Mp = Mp; w.Np = Np; w.m_valid = m_valid; w.n_valid = n_valid;
  w.X[0] = reinterpret_cast<const bf16*>(X); w.ldx[0] = ldx;
  w.Y[0] = reinterpret_cast<const bf16*>(Y); w.ldy[0] = ldy;
  w.n_pairs = 1; w.dW = dW; w.ldw = ldw; w.part = ws;
  return launch_tc_wgrad(w, as_stream(s));
}
}
