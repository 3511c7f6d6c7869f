// Fused SDF training chains (sdf_fused.cu): argument block shared with the host code in mlp_bf16.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "mlp_shape.cuh"
#include "tc_gemm.cuh"

namespace cope {

enum FzMode : int {
  FZ_FWD = 0,   // PE -> value pass (stores H_1..H_top) -> sdf / feature -> reverse sweep (stores delta_l) -> ge0 / ge1
  FZ_TAN = 1,   // tangent pass of the double backward: T_{l+1}, zb2_l
  FZ_ADJ = 2,   // adjoint pass: zb_l (optionally with the second-order term zb2_l), eb0 / eb1
  FZ_ADJ1 = 3,  // the same pass launched without zb2 (has_d == 0): picked by launch_sdf_fused, deeper weight ring
};

constexpr int kFzMaxJobs = 20;
// one tcgen05 accumulation  acc[128 x Np] (+)= A_tile[128 x Kp] * Wpacked(w_off)^T ; the MMA thread and the weight
// producer walk the same list
struct FzJob {
  uint32_t w_off;      // element offset of the packed [Kp/8][Np][8] block inside `wp`
  uint16_t Np, Kp;
  uint8_t acc;         // TMEM accumulator buffer (0 / 1)
  uint8_t wait_a;      // 0: A tile already waited for, 1: wait per 64-column panel (a_ready), 2: wait for the TMA-loaded tile (a_init)
  uint8_t commit;      // 0: none, 1 / 2: arrive on acc_full[commit-1] once this job's MMAs have completed
  uint8_t pad;
};

struct FzArgs {
  int64_t P;
  const float* x;            // [P x 4] points (FWD: PE, TAN: PE Jacobian)
  const float* g;            // TAN: upstream gradient of the SDF gradient [P x 4]
  const float* Wflat;        // fp32 flat parameters (biases, row 0 of the last layer)
  const __nv_bfloat16* wp;   // packed bf16 weights
  FzJob jobs[kFzMaxJobs];
  int n_jobs;
  int n_lin, skip, skw, pe_w, d_in, L;
  int64_t b_off[COPE_MAX_LIN];
  int64_t w_top_off;         // float offset of row 0 of the last layer inside Wflat
  // FWD
  float* sdf; int sdf_ld; int has_feat;
  int infer;                 // FWD: no backward will follow: the reverse-sweep deltas are not stored
  int value_only;            // FWD: value pass only (SDFNetwork.sdf with gradients): no reverse sweep, no delta / ge0 / ge1
  float* ge0; float* ge1;    // [P x 64] fp32 gradients w.r.t. the PE (layer 0 / skip layer)
  // ADJ
  const float* d_sdf; int d_sdf_ld;
  float* eb0; float* eb1;    // [P x 64] fp32 (want_e)
  int has_d, store_out, want_e;
  long long* dbg;            // optional timeline buffer (profiling builds): CTA 0 stamps clock64() per role
};

// 3-D tensor maps (64 x 128 x 1 boxes, 128B swizzle) over bf16 [layers][P][ld] buffers
struct FzMaps {
  CUtensorMap in0;    // FWD: pe [P x 64] (store)   TAN: t0 [P x 64] (store)   ADJ: d_feat [P x 256] (load)
  CUtensorMap H;      // saved activations H_1..H_top: layer index l-1
  CUtensorMap D;      // reverse-sweep deltas delta_0..delta_{top-1}
  CUtensorMap out;    // FWD: feature slot of the colour input   TAN: T_1..T_top   ADJ: zb_0..zb_{top-1}
  CUtensorMap Z2;     // zb2_0..zb2_{top-1} (TAN: store, ADJ: load)
};

// bf16 [layers][rows][ld] (cols of them addressed), layer stride in elements
int make_tmap3(const __nv_bfloat16* ptr, uint64_t cols, uint64_t rows, uint64_t layers, uint64_t ld, uint64_t layer_stride,
               CUtensorMap* out);
bool sdf_fused_supported(const MlpShape& m);
int launch_sdf_fused(int mode, const FzArgs& a, const FzMaps& maps, cudaStream_t s);

}  // namespace cope

// ---------------------------------------------------------------------------------------------------------------------
// Fused colour-network chains (color_fused.cu): same engine, 5 activation panels for the 320-wide first layer
namespace cope {

enum CzMode : int {
  CZ_FWD = 0,   // [feat | x | PE(dirs) | normals | x_lo] -> 4 ReLU layers (TMA-store h_1..h_4) -> sigmoid -> rgb
  CZ_BWD = 1,   // dz_top -> ReLU-masked adjoints dz_3..dz_0 (TMA-stored for the weight gradients) -> d_feat (bf16) + d(x, PE, normals)
};

struct CzArgs {
  int64_t P;
  const float* Wflat;
  const __nv_bfloat16* wp;
  FzJob jobs[kFzMaxJobs];
  int n_jobs;
  int n_lin, d_out;
  int64_t b_off[COPE_MAX_LIN];
  // FWD: the 64-column tail of the input [x_hi(4) | PE_Lv(dirs) | normals(4) | x_lo(4) | 0] is built in the kernel
  const float* x; const float* dirs; int dirs_group; int Lv; const float* normals;
  float* rgb; float* rgb_saved;
  int infer;                 // FWD: no backward will follow: neither the input tail nor the hidden activations are stored
  // BWD
  const float* d_rgb; const float* rgb_in;
  float* rest;               // [P x 64] fp32: gradient w.r.t. the input tail, or null
  int want_dfeat;
  long long* dbg;
};

struct CzMaps {
  CUtensorMap feat;     // FWD: feature block of the colour input [P x 256] (load)     BWD: d_feat [P x 256] (store)
  CUtensorMap tail;     // FWD: tail block of the colour input [P x 64] (store)        BWD: dz_top [P x 128] (store)
  CUtensorMap H;        // hidden activations h_1..h_4, layer index l-1 (FWD store, BWD load)
  CUtensorMap DZ;       // BWD: adjoints dz_0..dz_3 (store)
};

bool color_fused_supported(const MlpShape& m, int d_feat, int rest_cols);
int launch_color_fused(int mode, const CzArgs& a, const CzMaps& maps, cudaStream_t s);

}  // namespace cope
