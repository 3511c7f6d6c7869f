// bf16 tcgen05 GEMMs for the NeuS MLPs (sm_100a).
//
// tc_gemm_kernel   C[M x N] = epi(A[M x K] * W[N x K]^T): persistent CTAs, one 128-row tile at a time.
//   * W (<= 160 KB, pre-packed in UMMA core-matrix order) is bulk-TMA'd into shared memory ONCE per CTA and
//     stays resident; A streams through a 4-stage cp.async ring (128 x 64 bf16 per stage);
//   * one elected thread issues tcgen05.mma (M=128, N<=256, K=16) into one of two TMEM accumulators, so the
//     epilogue of tile i overlaps the MMAs of tile i+1;
//   * 8 epilogue warps pull the accumulator with tcgen05.ld and apply the fused NeuS epilogues.
// tc_wgrad_kernel  dW[Mp x Np] += X^T Y with K = points (both operands MN-major), fp32 accumulation in TMEM over
//   the CTA's slice of points, then one pass of fp32 atomics.
//
// Shared-memory operand layout: no swizzle, 8 x 16 B core matrices (see tc_common.cuh::smem_desc).
#include <algorithm>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace cope {
using namespace tc;

constexpr int kTcThreads = 416;          // warps 0-7 epilogue, 8-11 producers, 12 MMA + TMEM owner
constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;
constexpr int kAStageBytes = kTileM * kChunkK * 2;   // 16 KB

struct Load16 { uint4 a, b; };
__device__ __forceinline__ void load16_bf16(const bf16* p, float (&f)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[2 * i] = bf16_lo(w[i]); f[2 * i + 1] = bf16_hi(w[i]); }
}
__device__ __forceinline__ void store16(void* base, int64_t elem_off, int is_f32, const float (&v)[16], int n_ok) {
  if (is_f32) {
    float* o = reinterpret_cast<float*>(base) + elem_off;
    if (n_ok >= 16 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < n_ok) o[i] = v[i];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(base) + elem_off;
    if (n_ok >= 16 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      uint4 a, b;
      a.x = pack_bf16(v[0], v[1]); a.y = pack_bf16(v[2], v[3]); a.z = pack_bf16(v[4], v[5]); a.w = pack_bf16(v[6], v[7]);
      b.x = pack_bf16(v[8], v[9]); b.y = pack_bf16(v[10], v[11]); b.z = pack_bf16(v[12], v[13]); b.w = pack_bf16(v[14], v[15]);
      reinterpret_cast<uint4*>(o)[0] = a;
      reinterpret_cast<uint4*>(o)[1] = b;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (i < n_ok) o[i] = __float2bfloat16(v[i]);
    }
  }
}

__device__ __forceinline__ float sp_from_h(float h, float hscale) {
  // softplus'(z) = 1 - exp(-100 softplus(z))
  return 1.0f - __expf(-kSoftplusBeta * h * hscale);
}

// fused epilogue for 16 consecutive columns [n0, n0+16) of row m
__device__ __forceinline__ void tc_epilogue16(const TcArgs& a, int64_t m, int n0, float (&acc)[16]) {
  float o1[16], o2[16];
  float hv[16], dv[16];
  const bool needH = a.epi == TC_MUL_SIGP || a.epi == TC_TANGENT || a.epi == TC_BWD || a.epi == TC_RELU_MASK;
  if (needH && n0 < a.nsplit) load16_bf16(a.H + m * a.ldh + n0, hv);
  const bool needD = (a.epi == TC_TANGENT) || (a.epi == TC_BWD && a.D != nullptr);
  if (needD && n0 < a.nsplit) load16_bf16(a.D + m * a.ldd + n0, dv);
  const float r1 = a.r1 ? a.r1[m * a.r1_ld] : 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int n = n0 + i;
    float v = acc[i];
    if (a.bias && n < a.n_valid) v += a.bias[n];
    if (a.r1) v += r1 * a.r1w[n];
    o2[i] = 0.0f;
    switch (a.epi) {
      case TC_STORE: v *= a.alpha; break;
      case TC_BIAS_SOFTPLUS: v = a.alpha * softplus100(v); break;
      case TC_BIAS_RELU: v = fmaxf(v, 0.0f); break;
      case TC_BIAS_SIGMOID: v = sigmoidf_(v); break;
      case TC_MUL_SIGP:
        v = (n < a.nsplit) ? a.alpha * v * sp_from_h(hv[i], a.hscale) : a.alpha * v;
        break;
      case TC_TANGENT: {
        float sp = sp_from_h(hv[i], a.hscale);
        o2[i] = v * dv[i] * (kSoftplusBeta * (1.0f - sp));
        v = a.alpha * v * sp;
      } break;
      case TC_BWD:
        if (n < a.nsplit) {
          v = a.alpha * v * sp_from_h(hv[i], a.hscale);
          if (a.D) v += dv[i];
        } else {
          v = a.alpha * v;
        }
        break;
      case TC_RELU_MASK: v = hv[i] > 0.0f ? v : 0.0f; break;
    }
    o1[i] = v;
  }
  if (a.epi == TC_TANGENT) {
    store16(a.out, m * a.ldo + n0, a.out_f32, o1, a.n_valid - n0);
    store16(a.out2, m * a.ldo2 + n0, a.out2_f32, o2, a.n_valid - n0);
    return;
  }
  // split outputs: columns < nsplit -> out, columns >= nsplit -> out2 (shifted)
  if (n0 + 16 <= a.nsplit) {
    store16(a.out, m * a.ldo + n0, a.out_f32, o1, a.n_valid - n0);
  } else if (n0 >= a.nsplit) {
    if (a.out2) store16(a.out2, m * a.ldo2 + (n0 - a.nsplit), a.out2_f32, o1, a.n2_valid - (n0 - a.nsplit));
  } else {   // straddling chunk
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n < a.nsplit) {
        if (n < a.n_valid) {
          if (a.out_f32) reinterpret_cast<float*>(a.out)[m * a.ldo + n] = o1[i];
          else reinterpret_cast<bf16*>(a.out)[m * a.ldo + n] = __float2bfloat16(o1[i]);
        }
      } else if (a.out2 && (n - a.nsplit) < a.n2_valid) {
        if (a.out2_f32) reinterpret_cast<float*>(a.out2)[m * a.ldo2 + (n - a.nsplit)] = o1[i];
        else reinterpret_cast<bf16*>(a.out2)[m * a.ldo2 + (n - a.nsplit)] = __float2bfloat16(o1[i]);
      }
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbytes = (uint32_t)a.K * a.N * 2;
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((wbytes + 1023) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kStages * kAStageBytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = bars + 1 + kStages;
  uint64_t* acc_full = bars + 1 + 2 * kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(a_full + s, 128); mbar_init(a_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, 8); }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntiles = (a.M + kTileM - 1) / kTileM;
  const int nkc = a.K / kChunkK;

  if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ producers: weights once, A ring
    const int p = threadIdx.x - 256;
    if (p == 0) {
      mbar_arrive_expect_tx(w_full, wbytes);
      const uint32_t chunk = 32768;
      for (uint32_t off = 0; off < wbytes; off += chunk)
        bulk_g2s(sW + off, reinterpret_cast<const uint8_t*>(a.Bp) + off, min(chunk, wbytes - off), w_full);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t row0 = (int64_t)tile * kTileM;
      for (int kc = 0; kc < nkc; ++kc) {
        mbar_wait(a_empty + stage, phase ^ 1);
        const uint32_t sbase = smem_u32(sA + stage * kAStageBytes);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = i * 128 + p;                 // 16-byte chunk id == linear smem position
          const int rg = c >> 6, within = c & 63;
          const int k8 = within >> 3, r8 = within & 7;
          const int64_t row = row0 + rg * 8 + r8;
          const bool ok = row < a.M;
          const bf16* src = a.A + (ok ? row : 0) * a.lda + kc * kChunkK + k8 * 8;
          cp_async16(sbase + c * 16, src, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(a_full + stage);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(w_full, 0);
      const uint32_t idesc = idesc_bf16(kTileM, a.N, 0, 0);
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
          mbar_wait(a_full + stage, phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sAa = smem_u32(sA + stage * kAStageBytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks) {
            const uint64_t ad = smem_desc(sAa + ks * 256, 128, 1024);
            const uint64_t bd = smem_desc(sWa + (uint32_t)(kc * 8 + ks * 2) * b_lbo, b_lbo, 128);
            umma_bf16(d_tmem, ad, bd, idesc, (kc | ks) != 0);
          }
          umma_commit(a_empty + stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + acc);
        acc ^= 1;
        if (acc == 0) accp ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 0..7
    const int q = warp & 3, half = warp >> 2;
    const int nch = a.N / 16;
    const int c_begin = half ? (nch + 1) / 2 : 0, c_end = half ? nch : (nch + 1) / 2;
    uint32_t accp = 0;
    int acc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(acc_full + acc, accp);
      tc_fence_after();
      const int64_t m = (int64_t)tile * kTileM + q * 32 + lane;
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      for (int c = c_begin; c < c_end; ++c) {
        float v[16];
        tmem_ld16(taddr + c * 16, v);
        if (m < a.M) tc_epilogue16(a, m, c * 16, v);
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
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

static size_t tc_gemm_smem(int N, int K) {
  size_t w = ((size_t)K * N * 2 + 1023) & ~(size_t)1023;
  return w + kStages * kAStageBytes + 256;
}

int launch_tc_gemm(const TcArgs& a, cudaStream_t s) {
  if (a.M <= 0) return 0;
  COPE_REQUIRE(a.N % 16 == 0 && a.N >= 16 && a.N <= 256, "tc_gemm: N=%d must be a multiple of 16 in [16,256]", a.N);
  COPE_REQUIRE(a.K % kChunkK == 0 && a.K >= kChunkK && a.K <= 320, "tc_gemm: K=%d must be a multiple of 64 in [64,320]", a.K);
  COPE_REQUIRE(a.lda % 8 == 0 && a.lda >= a.K, "tc_gemm: lda=%d must be >= K and a multiple of 8", a.lda);
  static bool attr_set = false;
  const size_t smem = tc_gemm_smem(a.N, a.K);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    COPE_REQUIRE(e == cudaSuccess, "tc_gemm: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int sms = 148;
  const int ntiles = (a.M + kTileM - 1) / kTileM;
  tc_gemm_kernel<<<std::min(ntiles, sms), kTcThreads, smem, s>>>(a);
  COPE_CHECK_LAUNCH("tc_gemm");
  return 0;
}

// ================================================================================================ wgrad
constexpr int kWgStages = 3;
constexpr int kWgChunkP = 64;            // points per pipeline stage

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad_kernel(const TcWgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t xbytes = kWgChunkP * a.Mp * 2, ybytes = kWgChunkP * a.Np * 2;
  const uint32_t stage_bytes = xbytes + ybytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
  uint64_t* s_full = bars;
  uint64_t* s_empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(s_full + s, 128); mbar_init(s_empty + s, 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's slice of points (multiples of the stage size)
  const int64_t nchunks_total = (a.P + kWgChunkP - 1) / kWgChunkP;
  const int64_t per = (nchunks_total + gridDim.x - 1) / gridDim.x;
  const int64_t ch0 = (int64_t)blockIdx.x * per, ch1 = min(nchunks_total, ch0 + per);
  const int nmb = a.Mp / 128;
  const bool have_work = ch0 < ch1;

  if (warp >= 8 && warp < 12) {
    const int p = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int pr = 0; pr < a.n_pairs; ++pr) {
      for (int64_t ch = ch0; ch < ch1; ++ch) {
        mbar_wait(s_empty + stage, phase ^ 1);
        const uint32_t sx = smem_u32(smem + stage * stage_bytes), sy = sx + xbytes;
        const int64_t p0 = ch * kWgChunkP;
        // X tile: 64 points x Mp columns -> chunk c: p8 = c&7, j = (c>>3) % (Mp/8), pg = c / Mp
        const int xch = kWgChunkP * a.Mp / 8;
        for (int c = p; c < xch; c += 128) {
          const int p8 = c & 7, j = (c >> 3) % (a.Mp >> 3), pg = c / a.Mp;
          const int64_t pt = p0 + pg * 8 + p8;
          const bool ok = pt < a.P;
          cp_async16(sx + c * 16, a.X[pr] + (ok ? pt : 0) * a.ldx[pr] + j * 8, ok ? 16u : 0u);
        }
        const int ych = kWgChunkP * a.Np / 8;
        for (int c = p; c < ych; c += 128) {
          const int p8 = c & 7, j = (c >> 3) % (a.Np >> 3), pg = c / a.Np;
          const int64_t pt = p0 + pg * 8 + p8;
          const bool ok = pt < a.P;
          cp_async16(sy + c * 16, a.Y[pr] + (ok ? pt : 0) * a.ldy[pr] + j * 8, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(s_full + stage);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 12) {
    if (lane == 0 && have_work) {
      const uint32_t idesc = idesc_bf16(128, a.Np, 1, 1);
      const uint32_t x_lbo = (uint32_t)a.Mp * 16, y_lbo = (uint32_t)a.Np * 16;   // one 8-point slab
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int pr = 0; pr < a.n_pairs; ++pr) {
        for (int64_t ch = ch0; ch < ch1; ++ch) {
          mbar_wait(s_full + stage, phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + stage * stage_bytes), sy = sx + xbytes;
#pragma unroll
          for (int ks = 0; ks < kWgChunkP / 16; ++ks) {
            const uint64_t bd = smem_desc(sy + ks * 2 * y_lbo, y_lbo, 128);
            for (int mb = 0; mb < nmb; ++mb) {
              const uint64_t ad = smem_desc(sx + ks * 2 * x_lbo + mb * 2048, x_lbo, 128);
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
  } else if (have_work) {
    const int q = warp & 3, half = warp >> 2;
    const int nch = a.Np / 16;
    const int c_begin = half ? (nch + 1) / 2 : 0, c_end = half ? nch : (nch + 1) / 2;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int mb = 0; mb < nmb; ++mb) {
      const int m = mb * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + mb * 256 + ((uint32_t)(q * 32) << 16);
      for (int c = c_begin; c < c_end; ++c) {
        float v[16];
        tmem_ld16(taddr + c * 16, v);
        if (m < a.m_valid) {
          float* o = a.dW + (int64_t)m * a.ldw + c * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c * 16 + i < a.n_valid) atomicAdd(o + i, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 512);
}

int launch_tc_wgrad(const TcWgradArgs& a, cudaStream_t s) {
  if (a.P <= 0 || a.n_pairs <= 0) return 0;
  COPE_REQUIRE((a.Mp == 128 || a.Mp == 256) && a.Np % 16 == 0 && a.Np >= 16 && a.Np <= 256,
               "tc_wgrad: Mp=%d Np=%d unsupported", a.Mp, a.Np);
  for (int i = 0; i < a.n_pairs; ++i)
    COPE_REQUIRE(a.ldx[i] % 8 == 0 && a.ldy[i] % 8 == 0 && a.ldx[i] >= a.Mp && a.ldy[i] >= a.Np,
                 "tc_wgrad: operand %d leading dims (%d,%d) must cover the padded tile (%d,%d)", i, a.ldx[i], a.ldy[i], a.Mp, a.Np);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    COPE_REQUIRE(e == cudaSuccess, "tc_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t smem = (size_t)kWgStages * kWgChunkP * (a.Mp + a.Np) * 2 + 256;
  const int64_t nchunks = (a.P + kWgChunkP - 1) / kWgChunkP;
  const int grid = (int)std::min<int64_t>(148, std::max<int64_t>(1, nchunks / 4));
  tc_wgrad_kernel<<<grid, kTcThreads, smem, s>>>(a);
  COPE_CHECK_LAUNCH("tc_wgrad");
  return 0;
}

// ================================================================================================ packing
__device__ __forceinline__ int map_seg(const PackSeg* s, int n, int i) {
  for (int q = 0; q < n; ++q)
    if (i >= s[q].dst && i < s[q].dst + s[q].len) return s[q].src + (i - s[q].dst);
  return -1;
}
__global__ void tc_pack_kernel(const float* __restrict__ W, int ldw, const PackSpec spec, int Np, int Kp, int transposed,
                               bf16* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np * Kp) return;
  const int kk = i & 7, n = (i >> 3) % Np, kg = i / (8 * Np);
  const int sn = map_seg(spec.n, spec.nn, n), sk = map_seg(spec.k, spec.nk, kg * 8 + kk);
  float v = 0.0f;
  if (sn >= 0 && sk >= 0) v = transposed ? W[(int64_t)sk * ldw + sn] : W[(int64_t)sn * ldw + sk];
  out[i] = __float2bfloat16(v);
}

int launch_tc_pack(const float* W, int ldw, const PackSpec& spec, int Np, int Kp, int transposed, bf16* out, cudaStream_t s) {
  tc_pack_kernel<<<(Np * Kp + 255) / 256, 256, 0, s>>>(W, ldw, spec, Np, Kp, transposed, out);
  COPE_CHECK_LAUNCH("tc_pack");
  return 0;
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

int cope_tc_wgrad(int64_t P, int Mp, int Np, int m_valid, int n_valid, const void* X, int ldx, const void* Y, int ldy,
                  float* dW, int ldw, cope_stream_t s) {
  TcWgradArgs w{};
  w.P = P; w.Mp = Mp; w.Np = Np; w.m_valid = m_valid; w.n_valid = n_valid;
  w.X[0] = reinterpret_cast<const bf16*>(X); w.ldx[0] = ldx;
  w.Y[0] = reinterpret_cast<const bf16*>(Y); w.ldy[0] = ldy;
  w.n_pairs = 1; w.dW = dW; w.ldw = ldw;
  return launch_tc_wgrad(w, as_stream(s));
}
}
