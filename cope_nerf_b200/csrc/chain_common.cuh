// Building blocks shared by the fused chain kernels (sdf_fused.cu, color_fused.cu): warp roles, the swizzled activation
// panels, TMA helpers, the mbarrier block and the epilogue-side ring protocol.  Included inside an anonymous namespace.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "sdf_fused.cuh"
#include "tc_common.cuh"

namespace cope {
namespace chain {
using namespace tc;

constexpr int kEpiWarps = 16;
constexpr int kWProd = 16, kMma = 17, kStore = 18, kAuxW = 19;
constexpr int kThreads = 20 * 32;
constexpr int kPanel = 128 * 128;        // 128 rows x 64 bf16
constexpr int kWStage = 256 * 64 * 2;    // N = 256 x K = 64: one weight chunk per activation panel
constexpr int kMaxRing = 8;              // barrier slots reserved per ring
constexpr int kMaxPanels = 8;            // a_ready slots
constexpr int kBarBytes = 512;
// shared-memory carve-up: [A panels | weight ring | auxiliary ring | staging ring | bias | barriers]
template <int NPANELS, int KW, int KAUX, int KSTG, int KBIAS> struct ChainLay {
  static constexpr int kPanels = NPANELS, kW = KW, kAux = KAUX, kStg = KSTG;
  static constexpr int oA = 0;
  static constexpr int oW = oA + NPANELS * kPanel;
  static constexpr int oAux = oW + KW * kWStage;
  static constexpr int oStg = oAux + KAUX * kPanel;
  static constexpr int oBias = oStg + KSTG * kPanel;
  static constexpr int oBars = oBias + KBIAS;
  static constexpr int kSmem = oBars + kBarBytes;
  static_assert(kSmem <= 232448, "fused chain: shared memory budget");
  static_assert(KW <= kMaxRing && KAUX <= kMaxRing && NPANELS <= kMaxPanels, "ring too deep for the barrier block");
};
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// byte offset of 16-byte chunk c8 (8 bf16) of row r inside one 128B-swizzled panel
__device__ __forceinline__ uint32_t pan_off(int r, int c8) { return (uint32_t)r * 128 + (uint32_t)((c8 ^ (r & 7)) << 4); }
// scalar element k (0..255) of row r inside the 4-panel tile
__device__ __forceinline__ void put_elem(uint8_t* tile, int r, int k, float v) {
  *reinterpret_cast<bf16*>(tile + (k >> 6) * kPanel + pan_off(r, (k & 63) >> 3) + (k & 7) * 2) = __float2bfloat16(v);
}
__device__ __forceinline__ void write16(uint8_t* panel, int r, int part, const float (&v)[16]) {
  *reinterpret_cast<uint4*>(panel + pan_off(r, part * 2)) =
      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4*>(panel + pan_off(r, part * 2 + 1)) =
      make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}
struct Pk16 { uint32_t w[8]; };   // 16 bf16
__device__ __forceinline__ Pk16 read16(const uint8_t* panel, int r, int part) {
  const uint4 a = *reinterpret_cast<const uint4*>(panel + pan_off(r, part * 2));
  const uint4 b = *reinterpret_cast<const uint4*>(panel + pan_off(r, part * 2 + 1));
  Pk16 p;
  p.w[0] = a.x; p.w[1] = a.y; p.w[2] = a.z; p.w[3] = a.w; p.w[4] = b.x; p.w[5] = b.y; p.w[6] = b.z; p.w[7] = b.w;
  return p;
}
__device__ __forceinline__ float pk_get(const Pk16& p, int i) {
  const uint32_t w = p.w[i >> 1];
  return (i & 1) ? bf16_hi(w) : bf16_lo(w);
}

__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// timeline stamps (CTA 0 only, when a.dbg != nullptr): region `role` holds (tag << 48 | clock) entries
struct Stamp {
  long long* p; int n;
  __device__ __forceinline__ void init(long long* base, int role) { p = (base && blockIdx.x == 0) ? base + role * 4096 : nullptr; n = 1; }
  __device__ __forceinline__ void operator()(int tag) {
    if (p && n < 4096) { p[n++] = ((long long)tag << 48) | (clock64() & 0xFFFFFFFFFFFFll); p[0] = n; }
  }
};

// mbarrier wait for the single-thread roles (producers, MMA issuer, store issuer): they sit on the same schedulers as
// the epilogue warps, so a failed probe parks the thread (suspend-time hint) instead of re-issuing the probe at once
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
  }
  __trap();
}

struct Bars {
  uint64_t *w_full, *w_empty, *aux_full, *aux_empty, *a_ready, *acc_full, *stg_full, *stg_empty, *a_free, *a_init, *tile_done,
      *h_stored, *epi_done;
  uint32_t* tmem_slot;
  __device__ __forceinline__ void carve(uint64_t* bars) {
    w_full = bars; w_empty = bars + kMaxRing; aux_full = bars + 2 * kMaxRing; aux_empty = bars + 3 * kMaxRing;
    a_ready = bars + 4 * kMaxRing; acc_full = a_ready + kMaxPanels; stg_full = acc_full + 2; stg_empty = stg_full + 2;
    a_free = stg_empty + 2; a_init = a_free + 1; tile_done = a_free + 2; h_stored = a_free + 3; epi_done = a_free + 4;
    tmem_slot = reinterpret_cast<uint32_t*>(a_free + 5);
  }
  // called by one thread; npanels a_ready barriers, rings as configured
  __device__ __forceinline__ void init(int kw, int kaux, int npanels) {
    for (int s = 0; s < kw; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
    for (int s = 0; s < kaux; ++s) { mbar_init(aux_full + s, 1); mbar_init(aux_empty + s, kEpiWarps); }
    for (int j = 0; j < npanels; ++j) mbar_init(a_ready + j, kEpiWarps);
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(stg_full + s, kEpiWarps); mbar_init(stg_empty + s, 1); }
    mbar_init(a_free, 1); mbar_init(a_init, 1); mbar_init(tile_done, 1); mbar_init(h_stored, 1);
    mbar_init(epi_done, kEpiWarps);
    fence_barrier_init();
  }
};
static_assert((4 * kMaxRing + kMaxPanels + 2 + 2 + 2 + 5) * 8 + 8 <= kBarBytes, "barrier block too small");

// epilogue-side view of the rings
template <int kAuxRing, int kStgRing>
struct EpiCtx {
  uint8_t *sA, *sAux, *sStg;
  Bars B;
  int r, q, part, lane;
  uint32_t ev, auxc, stgc, accp;
  uint32_t tmem_base;

  __device__ __forceinline__ void begin_event() {          // about to overwrite A panels: the previous event's TMA stores
    if (ev > 0) mbar_wait(B.a_free, (ev - 1) & 1);         // must have finished reading them
    ++ev;
    st(510);
  }
  Stamp st;
  __device__ __forceinline__ void panel_done(int j) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(B.a_ready + j);
    st(600 + j);
  }
  __device__ __forceinline__ uint32_t wait_acc(int b) {
    mbar_wait(B.acc_full + b, (accp >> b) & 1);
    accp ^= 1u << b;
    tc_fence_after();
    st(500 + b);
    return tmem_base + b * 256 + ((uint32_t)(q * 32) << 16);
  }
  // wait for the next auxiliary panel and pull this thread's 16 values; the slot is handed back (aux_release) only after
  // the values have been CONSUMED: releasing right after the loads were issued let the refill overtake them
  __device__ __forceinline__ Pk16 aux_take() {
    const uint32_t slot = auxc % kAuxRing, par = (auxc / kAuxRing) & 1;
    mbar_wait(B.aux_full + slot, par);
    ++auxc;
    st(520);
    return read16(sAux + slot * kPanel, r, part);
  }
  __device__ __forceinline__ void aux_release(int n) {     // the n most recently taken slots
    __syncwarp();
    if (lane == 0)
      for (int k = n; k >= 1; --k) mbar_arrive(B.aux_empty + ((auxc - k) % kAuxRing));
  }
  __device__ __forceinline__ void stg_put(const float (&v)[16]) {
    constexpr uint32_t kS = kStgRing > 0 ? kStgRing : 1;
    const uint32_t slot = stgc % kS, par = (stgc / kS) & 1;
    mbar_wait(B.stg_empty + slot, par ^ 1);
    write16(sStg + slot * kPanel, r, part, v);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(B.stg_full + slot);
    ++stgc;
  }
};


}  // namespace chain
}  // namespace cope
