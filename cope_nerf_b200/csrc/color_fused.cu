// Fused colour-network chains for the bf16 tcgen05 path (RenderingNetwork 'idr': 291 -> 4 x 256 ReLU -> 3, sigmoid).
// Same engine as sdf_fused.cu (chain_common.cuh): one persistent CTA per SM, the activation tile resident in 128B-swizzled
// shared-memory panels, 32 KB weight chunks streamed through a bulk-TMA ring, tcgen05.mma into two TMEM accumulators,
// 16 epilogue warps, saved tiles moved by TMA.
//
//   CZ_FWD  A = [feat (4 panels, TMA-loaded from the slot the SDF chain stored them into) | tail panel built here:
//           x_hi, PE(dirs), normals, x_lo]; 4 x (bias + ReLU), each h_l TMA-stored for the backward; sigmoid -> rgb
//   CZ_BWD  dz_top = d_rgb rgb (1 - rgb); dz_{l-1} = (h_l > 0) (W_l^T dz_l), every dz_l TMA-stored for the weight
//           gradients; d_feat (bf16, into the SDF backward's slot) and the fp32 gradient of the input tail
#include <stdlib.h>

#include <algorithm>

#include "chain_common.cuh"
#include "sdf_fused.cuh"
#include "tc_common.cuh"

namespace cope {
using namespace tc;
using namespace chain;

namespace {

// ring depths (weight / auxiliary), measured at 131 072 points: forward 3 / 0: 112 us, 4 / 0: 109 us; backward 3 / 4: 153 us, 2 / 6: 178 us,
// 4 / 2: 149 us -- the weight ring is the one that has to be deep
template <int MODE> struct CCfg;
template <> struct CCfg<CZ_FWD> { static constexpr int kP = 5, kW = 4, kAux = 0, kStg = 0, kBias = COPE_MAX_LIN * 256 * 4; };
template <> struct CCfg<CZ_BWD> { static constexpr int kP = 4, kW = 4, kAux = 2, kStg = 0, kBias = 0; };
template <int MODE> using CLay = ChainLay<CCfg<MODE>::kP, CCfg<MODE>::kW, CCfg<MODE>::kAux, CCfg<MODE>::kStg, CCfg<MODE>::kBias>;

// element e of the 64-column input tail: [x_hi(4) | d, sin/cos(2^k d) (3 + 6 Lv) | normals(4) | x_lo(4) | 0].
// Select-by-compare instead of x[e] / d[q] / nrm[..]: a run-time index puts the arrays in local memory, and with the shared-memory
// carve-out at its maximum there is no L1 behind it.  sin / cos through the MUFU (|2^k d| <= 2^(Lv-1): abs. error ~1e-6, against
// the 4e-3 of the bf16 rounding that follows): the library sinf + cosf made this the longest serial step of a tile, 16 elements x
// ~160 instructions per thread before the first MMA can start.
__device__ __forceinline__ float csel3(const float (&v)[3], int i) { return i == 0 ? v[0] : i == 1 ? v[1] : v[2]; }
__device__ __forceinline__ float csel4(const float (&v)[4], int i) { return i == 0 ? v[0] : i == 1 ? v[1] : i == 2 ? v[2] : v[3]; }
__device__ __forceinline__ float tail_elem(const float (&x)[4], const float (&d)[3], const float (&nrm)[4], int Lv, int e) {
  const int pe_w = 3 * (1 + 2 * Lv);
  if (e < 4) return __bfloat162float(__float2bfloat16(csel4(x, e)));
  if (e < 4 + pe_w) {
    const int q = e - 4;
    if (q < 3) return csel3(d, q);
    const int blk = (q - 3) / 3, dd = (q - 3) - blk * 3;
    const float ang = csel3(d, dd) * (float)(1 << (blk >> 1));
    return (blk & 1) ? __cosf(ang) : __sinf(ang);
  }
  if (e < 8 + pe_w) return csel4(nrm, e - 4 - pe_w);
  if (e < 12 + pe_w) { const float xv = csel4(x, e - 8 - pe_w); return xv - __bfloat162float(__float2bfloat16(xv)); }
  return 0.0f;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) color_fused_kernel(const __grid_constant__ CzArgs a, const __grid_constant__ CzMaps tm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = CLay<MODE>;
  constexpr int kWRing = L::kW, kAuxRing = L::kAux > 0 ? L::kAux : 1, kPanels = L::kPanels;
  uint8_t* sA = smem + L::oA;
  uint8_t* sW = smem + L::oW;
  uint8_t* sAux = smem + L::oAux;
  float* sBias = reinterpret_cast<float*>(smem + L::oBias);
  Bars B;
  B.carve(reinterpret_cast<uint64_t*>(smem + L::oBars));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) B.init(kWRing, L::kAux, kPanels);
  if (MODE == CZ_FWD) {
    for (int i = threadIdx.x; i < a.n_lin * 256; i += kThreads) {
      const int l = i >> 8, n = i & 255;
      const int nout = l == a.n_lin - 1 ? a.d_out : 256;
      sBias[i] = n < nout ? a.Wflat[a.b_off[l] + n] : 0.0f;
    }
  }
  if (warp == kMma) tmem_alloc(B.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;
  const int top = a.n_lin - 1;            // 4
  const int tile_first = blockIdx.x, tile_step = gridDim.x;
  const int ntiles = (int)((a.P + 127) / 128);

  if (warp == kWProd) {
    // ================================================================== weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        for (int jb = 0; jb < a.n_jobs; ++jb) {
          const FzJob J = a.jobs[jb];
          const int nch = J.Kp >> 6;
          const uint32_t cbytes = (uint32_t)J.Np * 128;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wp + J.w_off);
          for (int c = 0; c < nch; ++c) {
            mbar_wait_park(B.w_empty + stage, phase ^ 1);
            mbar_arrive_expect_tx(B.w_full + stage, cbytes);
            bulk_g2s(sW + stage * kWStage, src + (size_t)c * cbytes, cbytes, B.w_full + stage);
            if (++stage == kWRing) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMma) {
    // ================================================================== MMA issuer
    // FzJob::wait_a: 0 none, 1 a_ready[c] per chunk, 3 a_init (feature panels) then a_ready[4] for the tail chunk;
    // FzJob::pad: bit mask of panels whose a_ready phase this job must drain afterwards (written, but not multiplied)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aph = 0, initph = 0;
      const uint64_t adesc0 = smem_desc_sw128(smem_u32(sA), 16, 1024);
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        for (int jb = 0; jb < a.n_jobs; ++jb) {
          const FzJob J = a.jobs[jb];
          const int nch = J.Kp >> 6;
          const uint32_t idesc = idesc_bf16(128, J.Np, 0, 0);
          const uint32_t b_lbo = (uint32_t)J.Np * 16;
          const uint32_t d_tmem = tmem_base + J.acc * 256;
          const uint64_t bdesc0 = smem_desc(smem_u32(sW), b_lbo, 128);
          const uint32_t b_kstep = (2 * b_lbo) >> 4;
          if (J.wait_a == 3) { mbar_wait_park(B.a_init, initph); initph ^= 1; }
          for (int c = 0; c < nch; ++c) {
            if (J.wait_a == 1 || (J.wait_a == 3 && c == 4)) {
              mbar_wait_park(B.a_ready + c, (aph >> c) & 1);
              aph ^= 1u << c;
            }
            mbar_wait_park(B.w_full + stage, phase);
            tc_fence_after();
            const uint64_t ad = adesc0 + (uint64_t)(c * (kPanel >> 4));
            const uint64_t bd = bdesc0 + (uint64_t)(stage * (kWStage >> 4));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_bf16(d_tmem, ad + ks * 2, bd + ks * b_kstep, idesc, (c | ks) != 0);
            umma_commit(B.w_empty + stage);
            if (++stage == kWRing) { stage = 0; phase ^= 1; }
          }
          if (J.commit) umma_commit(B.acc_full + (J.commit - 1));
          for (int j = 0; j < kPanels; ++j)
            if (J.pad & (1u << j)) { mbar_wait_park(B.a_ready + j, (aph >> j) & 1); aph ^= 1u << j; }
        }
        if (MODE == CZ_FWD) umma_commit(B.tile_done);
      }
    }
  } else if (warp == kStore) {
    // ================================================================== TMA-store issuer
    if (lane == 0) {
      uint32_t aph = 0;
      auto wait_panel = [&](int j) { mbar_wait_park(B.a_ready + j, (aph >> j) & 1); aph ^= 1u << j; };
      auto store_tile = [&](const CUtensorMap* map, int row0, int layer, bool do_store) {
        for (int j = 0; j < 4; ++j) {
          wait_panel(j);
          if (do_store) tma_store_3d(map, sA + j * kPanel, j * 64, row0, layer);
        }
        if (do_store) { bulk_commit(); bulk_wait_read0(); }
        mbar_arrive(B.a_free);
      };
      for (int tile = tile_first; tile < ntiles; tile += tile_step) {
        const int row0 = tile * 128;
        if (MODE == CZ_FWD) {
          wait_panel(4);                                                          // input tail -> saved colour input
          if (!a.infer) {
            tma_store_3d(&tm.tail, sA + 4 * kPanel, 0, row0, 0);
            bulk_commit(); bulk_wait_read0();
          }
          mbar_arrive(B.a_free);
          for (int l = 0; l < top; ++l) store_tile(&tm.H, row0, l, !a.infer);     // h_{l+1}
        } else {
          wait_panel(0); wait_panel(1);                                           // dz_top, zero-padded to 128 columns
          tma_store_3d(&tm.tail, sA, 0, row0, 0);
          tma_store_3d(&tm.tail, sA + kPanel, 64, row0, 0);
          bulk_commit(); bulk_wait_read0();
          mbar_arrive(B.a_free);
          for (int l = top - 1; l >= 0; --l) store_tile(&tm.DZ, row0, l, true);   // dz_l
          store_tile(&tm.feat, row0, 0, a.want_dfeat != 0);                       // d_feat
        }
      }
      bulk_wait0();
    }
  } else if (warp == kAuxW) {
    // ================================================================== auxiliary producer
    if (lane == 0) {
      uint32_t auxc = 0, t_local = 0;
      for (int tile = tile_first; tile < ntiles; tile += tile_step, ++t_local) {
        const int row0 = tile * 128;
        if (MODE == CZ_FWD) {
          // feature panels of this tile into A panels 0..3: the previous tile's MMAs, epilogue and stores are done with them
          if (t_local > 0) {
            mbar_wait_park(B.tile_done, (t_local - 1) & 1);
            mbar_wait_park(B.epi_done, (t_local - 1) & 1);
            const uint32_t evs = t_local * (uint32_t)(top + 1);                   // A-write events so far
            mbar_wait_park(B.a_free, (evs - 1) & 1);
          }
          mbar_arrive_expect_tx(B.a_init, 4 * kPanel);
          for (int j = 0; j < 4; ++j) tma_load_3d(sA + j * kPanel, &tm.feat, j * 64, row0, 0, B.a_init);
        } else {
          for (int l = top; l >= 1; --l)
            for (int j = 0; j < 4; ++j) {                                         // h_l
              const uint32_t slot = auxc % kAuxRing, par = (auxc / kAuxRing) & 1;
              mbar_wait_park(B.aux_empty + slot, par ^ 1);
              mbar_arrive_expect_tx(B.aux_full + slot, kPanel);
              tma_load_3d(sAux + slot * kPanel, &tm.H, j * 64, row0, l - 1, B.aux_full + slot);
              ++auxc;
            }
        }
      }
    }
  } else {
    // ================================================================== epilogue warps 0..15
    EpiCtx<kAuxRing, 0> E;
    E.sA = sA; E.sAux = sAux; E.sStg = nullptr; E.B = B;
    E.q = warp & 3; E.part = warp >> 2; E.lane = lane; E.r = E.q * 32 + lane;
    E.ev = 0; E.auxc = 0; E.stgc = 0; E.accp = 0; E.tmem_base = tmem_base;
    E.st.init(nullptr, 1);
    const int r = E.r, part = E.part;
    for (int tile = tile_first; tile < ntiles; tile += tile_step) {
      const int64_t m = (int64_t)tile * 128 + r;
      const bool ok = m < a.P;
      const int64_t mm = ok ? m : 0;
      if (MODE == CZ_FWD) {
        // ---------------- input tail (panel 4): this thread's 16 columns
        {
          float x[4] = {0, 0, 0, 0}, d[3] = {0, 0, 0}, nrm[4] = {0, 0, 0, 0};
          if (ok) {
            const float4 t = *reinterpret_cast<const float4*>(a.x + mm * 4);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            const float4 g = *reinterpret_cast<const float4*>(a.normals + mm * 4);
            nrm[0] = g.x; nrm[1] = g.y; nrm[2] = g.z; nrm[3] = g.w;
            const float* dv = a.dirs + (mm / a.dirs_group) * 3;
            d[0] = dv[0]; d[1] = dv[1]; d[2] = dv[2];
          }
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = tail_elem(x, d, nrm, a.Lv, part * 16 + i);
          E.begin_event();
          write16(sA + 4 * kPanel, r, part, v);
          E.panel_done(4);
        }
        // ---------------- hidden layers: h_{l+1} = relu(W_l h_l + b_l)
        for (int l = 0; l < top; ++l) {
          const uint32_t taddr = E.wait_acc(l & 1);
          E.begin_event();
          const float* bias = sBias + l * 256;
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int n0 = j * 64 + part * 16;
            float v[16], bz[16];
            tmem_ld16(taddr + n0, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(bz + 4 * i) = *reinterpret_cast<const float4*>(bias + n0 + 4 * i);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + bz[i], 0.0f);
            write16(sA + j * kPanel, r, part, v);
            E.panel_done(j);
          }
          tc_fence_before();
        }
        // ---------------- output layer: rgb = sigmoid(W_top h_top + b_top)
        {
          const uint32_t taddr = E.wait_acc(top & 1);
          if (part == 0) {
            float v[16];
            tmem_ld16(taddr, v);
            if (ok) {
#pragma unroll
              for (int k = 0; k < 16; ++k) {              // unrolled with a guard: v[] stays in registers
                if (k < a.d_out) {
                  const float o = __fdividef(1.0f, 1.0f + ex2((v[k] + sBias[top * 256 + k]) * -1.4426950408889634f));
                  a.rgb[m * a.d_out + k] = o;
                  if (a.rgb_saved) a.rgb_saved[m * a.d_out + k] = o;
                }
              }
            }
          }
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(B.epi_done);
      } else {
        // ---------------- dz_top = d_rgb * rgb * (1 - rgb) into panel 0 (columns >= d_out and panel 1: zero)
        {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.0f;
          E.begin_event();
          write16(sA + kPanel, r, part, v);
          if (part == 0 && ok) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {                // unrolled with a guard: v[] stays in registers
              if (k < a.d_out) {
                const float o = a.rgb_in[m * a.d_out + k];
                v[k] = a.d_rgb[m * a.d_out + k] * o * (1.0f - o);
              }
            }
          }
          write16(sA, r, part, v);
          E.panel_done(0);
          E.panel_done(1);
        }
        // ---------------- dz_{l-1} = (h_l > 0) ? W_l^T dz_l : 0, l = top .. 1
        for (int l = top; l >= 1; --l) {
          const int s = top - l;
          const uint32_t taddr = E.wait_acc(s & 1);
          E.begin_event();
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int n0 = j * 64 + part * 16;
            const Pk16 hp = E.aux_take();
            float v[16];
            tmem_ld16(taddr + n0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t hb = (i & 1) ? (hp.w[i >> 1] >> 16) : (hp.w[i >> 1] & 0xFFFFu);
              v[i] = (hb & 0x7FFFu) ? v[i] : 0.0f;        // h = relu(.) >= 0: non-zero bits <=> h > 0
            }
            E.aux_release(1);
            write16(sA + j * kPanel, r, part, v);
            E.panel_done(j);
          }
          tc_fence_before();
        }
        // ---------------- layer 0: d_feat (acc 0, bf16 through the A panels) and the input-tail gradient (acc 1, fp32)
        {
          const uint32_t taddr = E.wait_acc(0);
          E.begin_event();
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int n0 = j * 64 + part * 16;
            float v[16];
            tmem_ld16(taddr + n0, v);
            write16(sA + j * kPanel, r, part, v);
            E.panel_done(j);
          }
          if (a.rest) {
            float v[16];
            tmem_ld16(tmem_base + 256 + ((uint32_t)(E.q * 32) << 16) + part * 16, v);
            if (ok) {
              float4* o = reinterpret_cast<float4*>(a.rest + m * 64 + part * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
          tc_fence_before();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMma) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool color_fused_supported(const MlpShape& m, int d_feat, int rest_cols) {
  if (m.n_lin != 5 || m.skip >= 0 || d_feat != 256 || rest_cols + 4 > 64 || m.d_out > 16) return false;
  for (int l = 1; l < m.n_lin; ++l)
    if (m.in[l] != 256) return false;
  for (int l = 0; l < m.n_lin; ++l)
    if (m.b_off[l] % 4 != 0) return false;
  return true;
}

template <int MODE>
static int launch_c(const CzArgs& a, const CzMaps& maps, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(color_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CLay<MODE>::kSmem);
    COPE_REQUIRE(e == cudaSuccess, "color_fused: cannot raise dynamic shared memory to %d: %s", CLay<MODE>::kSmem, cudaGetErrorString(e));
    attr_set = true;
  }
  const int ntiles = (int)((a.P + 127) / 128);
  color_fused_kernel<MODE><<<std::min(ntiles, 148), kThreads, CLay<MODE>::kSmem, s>>>(a, maps);
  COPE_CHECK_LAUNCH("color_fused");
  return 0;
}

int launch_color_fused(int mode, const CzArgs& a, const CzMaps& maps, cudaStream_t s) {
  if (a.P <= 0) return 0;
  COPE_REQUIRE(a.n_jobs > 0 && a.n_jobs <= kFzMaxJobs, "color_fused: bad job list (%d)", a.n_jobs);
  if (mode == CZ_FWD) return launch_c<CZ_FWD>(a, maps, s);
  if (mode == CZ_BWD) return launch_c<CZ_BWD>(a, maps, s);
  COPE_REQUIRE(false, "color_fused: unknown mode %d", mode);
}

}  // namespace cope
