// Shapes / flat-parameter layout shared by the fp32 and bf16 MLP paths.
#pragma once
#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"

namespace cope {

// ------------------------------------------------------------------------------------------- layouts
struct MlpShape {
  int n_lin, d_in, L, pe_w, skip, ldh, d_out;
  int act; float act_slope;
  int64_t w_off[COPE_MAX_LIN], b_off[COPE_MAX_LIN], n_flat;
  int in[COPE_MAX_LIN], out[COPE_MAX_LIN];
};

inline int make_shape(const cope_mlp_desc* d, MlpShape* s) {
  COPE_REQUIRE(d && d->n_lin >= 2 && d->n_lin <= COPE_MAX_LIN, "mlp desc: n_lin out of range");
  s->n_lin = d->n_lin; s->d_in = d->d_in; s->L = d->multires; s->skip = d->skip_layer;
  COPE_REQUIRE(d->activation == COPE_ACT_SOFTPLUS100 || d->activation == COPE_ACT_LEAKY_RELU, "mlp desc: unknown activation %d", d->activation);
  s->act = d->activation; s->act_slope = d->act_param;
  s->pe_w = d->d_in * (1 + 2 * d->multires);
  int64_t off = 0;
  int ldh = 0;
  for (int l = 0; l < d->n_lin; ++l) { s->in[l] = d->dims_in[l]; s->out[l] = d->dims_out[l]; }
  for (int l = 0; l < d->n_lin; ++l) {
    COPE_REQUIRE(s->in[l] > 0 && s->out[l] > 0, "mlp desc: bad dims at layer %d", l);
    s->w_off[l] = off; off += (int64_t)s->in[l] * s->out[l];
    s->b_off[l] = off; off += s->out[l];
    if (l > 0) ldh = std::max(ldh, s->in[l]);
    if (l + 1 < d->n_lin) {
      int expect = (l + 1 == s->skip) ? s->in[l + 1] - s->pe_w : s->in[l + 1];
      COPE_REQUIRE(s->out[l] == expect, "mlp desc: layer %d emits %d but layer %d expects %d", l, s->out[l], l + 1, expect);
    }
  }
  s->n_flat = off;
  s->ldh = (ldh + 3) / 4 * 4;
  s->d_out = s->out[d->n_lin - 1];
  return 0;
}


// bf16 tcgen05 implementations (mlp_bf16.cu); same contracts as the C entry points of include/cope_b200.h
int64_t sdf_saved_floats_bf16(const MlpShape& m, int64_t P, int with_grad);
int64_t sdf_ws_floats_bf16(const MlpShape& m, int64_t P);
bool sdf_infer_compact_bf16(const MlpShape& m);
int64_t sdf_query_ws_floats_bf16(const MlpShape& m, int64_t P);
int sdf_query_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, float* sdf_out, float* ws, cudaStream_t s,
                   bool ws_holds_pack = false, const __nv_bfloat16* wpx = nullptr);
int sdf_fwd_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, float* sdf, int sdf_ld, float* feat,
                 int feat_ld, float* grad, float* saved, float* ws, cudaStream_t s, __nv_bfloat16* feat_b16 = nullptr,
                 int feat_b16_ld = 0, bool infer = false, const __nv_bfloat16* wpx = nullptr);
int sdf_bwd_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, const float* saved, const float* d_sdf,
                 int d_sdf_ld, const float* d_feat, int d_feat_ld, const float* dgrad, float* dWflat, float* dx,
                 int dx_accumulate, float* ws, cudaStream_t s, bool d_feat_in_ws = false, const __nv_bfloat16* wpx = nullptr);
__nv_bfloat16* sdf_bwd_dfeat_slot(const MlpShape& m, int64_t P, float* ws, int* ld);
__nv_bfloat16* color_cin_slot(const MlpShape& m, int Lv, int64_t P, float* saved, int* ld);
int64_t color_saved_floats_bf16(const MlpShape& m, int64_t P);
int64_t color_ws_floats_bf16(const MlpShape& m, int64_t P);
int color_fwd_bf16(const MlpShape& m, const float* Wflat, const float* x, const float* dirs, int dirs_group, int Lv,
                   const float* normals, const float* feat, int feat_ld, int64_t P, float* rgb, float* saved, float* ws,
                   cudaStream_t s, bool feat_in_cin = false, bool infer = false, const __nv_bfloat16* wpx = nullptr);
int color_bwd_bf16(const MlpShape& m, const float* Wflat, const float* dirs, int dirs_group, int Lv, int64_t P,
                   const float* saved, const float* d_rgb, float* dWflat, float* dx, float* ddirs, float* dnormals,
                   float* dfeat, int dfeat_ld, float* ws, cudaStream_t s, __nv_bfloat16* dfeat_b16 = nullptr,
                   int dfeat_b16_ld = 0, const __nv_bfloat16* wpx = nullptr);
int64_t sdf_fwd_ws_floats_bf16(const MlpShape& m, int64_t P);
int64_t color_fwd_ws_floats_bf16(const MlpShape& m);
int64_t mlp_pack_elems_bf16(const MlpShape& m, int is_color, int Lv);
int mlp_pack_bf16(const MlpShape& m, int is_color, int Lv, const float* Wflat, __nv_bfloat16* wp, cudaStream_t s);
// COPE_FLAT_HAS_PACK: the packed bf16 weights follow the flat fp32 parameters (64-float aligned); nullptr when the flag is off
inline const __nv_bfloat16* flat_pack_ptr(const MlpShape& m, const float* Wflat, int prec) {
  return (prec & COPE_FLAT_HAS_PACK) ? reinterpret_cast<const __nv_bfloat16*>(Wflat + (m.n_flat + 63) / 64 * 64) : nullptr;
}

// small fp32 kernels of mlp_f32.cu that the bf16 path reuses
__global__ void pe_vjp_kernel(const float* __restrict__ x, int64_t P, int d, int L, const float* __restrict__ ge0, int ld0,
                              const float* __restrict__ ge1, int ld1, float* __restrict__ g, int ldg, int accumulate);

}  // namespace cope
