// render_core's MLP stage as one call: SDF value + analytic gradient, then the colour network, with the 256-wide
// feature vector (and its gradient on the way back) handed over in bf16 inside the colour net's input buffer on the
// tensor-core path — no fp32 round trip, no re-pack.  model/neus_renderer.py:352-358.
#include "mlp_shape.cuh"

using namespace cope;

namespace cope { int64_t sdf_bwd_eb_offset(const MlpShape& m, int64_t P); }

extern "C" {

/* test hook: float offset (inside the ws of cope_render_mlp_bwd) of the [P x 64] fp32 PE-gradient buffers eb0, eb1 */
int64_t cope_dbg_render_bwd_eb_offset(const cope_mlp_desc* sd, const cope_mlp_desc* cd, int64_t P) {
  MlpShape ms;
  if (make_shape(sd, &ms)) return -1;
  return cope_color_ws_floats(cd, P, COPE_PREC_BF16) + sdf_bwd_eb_offset(ms, P);
}

int64_t cope_render_mlp_ws_floats(const cope_mlp_desc* sd, const cope_mlp_desc* cd, int64_t P, int prec) {
  const int64_t a = cope_sdf_ws_floats(sd, P, prec), b = cope_color_ws_floats(cd, P, prec);
  if (a < 0 || b < 0) return -1;
  MlpShape ms;
  if (make_shape(sd, &ms)) return -1;
  return a + b + 2 * P * (int64_t)(ms.d_out - 1) + 1024;     // + fp32 feature / feature-gradient temporaries (strict path)
}

int cope_render_mlp_fwd(const cope_mlp_desc* sd, const float* sdfW, const cope_mlp_desc* cd, const float* colW, const float* x,
                        const float* dirs, int dirs_group, int Lv, int64_t P, float* sdf, float* grad, float* rgb,
                        float* sdf_saved, float* col_saved, float* ws, int prec, cope_stream_t s_) {
  MlpShape ms, mc;
  if (make_shape(sd, &ms) || make_shape(cd, &mc)) return -1;
  if (P <= 0) return 0;
  cudaStream_t s = as_stream(s_);
  if ((prec & 0xFF) == COPE_PREC_BF16) {
    int ld = 0;
    __nv_bfloat16* cin = color_cin_slot(mc, Lv, P, col_saved, &ld);
    COPE_REQUIRE(cin != nullptr, "render_mlp_fwd: colour network shape not supported on the bf16 path");
    if (int rc = sdf_fwd_bf16(ms, sdfW, x, P, sdf, 1, nullptr, 0, grad, sdf_saved, ws, s, cin, ld, false, flat_pack_ptr(ms, sdfW, prec)))
      return rc;
    return color_fwd_bf16(mc, colW, x, dirs, dirs_group, Lv, grad, nullptr, 0, P, rgb, col_saved, ws, s, true, false,
                          flat_pack_ptr(mc, colW, prec));
  }
  const int F = ms.d_out - 1;
  float* feat = ws;
  float* rest = ws + P * (int64_t)F;
  if (int rc = cope_sdf_fwd(sd, sdfW, x, P, sdf, 1, feat, F, grad, sdf_saved, rest, prec, s_)) return rc;
  return cope_color_fwd(cd, colW, x, dirs, dirs_group, Lv, grad, feat, F, P, rgb, col_saved, rest, prec, s_);
}

/* ---- inference: same outputs as cope_render_mlp_fwd, nothing kept for a backward pass (full-image rendering, eval.py) */
// saved-block size of the inference forward: H only (no delta stack) exactly when the bf16 forward runs on the fused chain;
// the layer-by-layer fallback (COPE_NO_FUSED, or a shape sdf_fused_supported rejects) still writes delta_l behind H
static int64_t infer_sdf_saved_floats(const cope_mlp_desc* sd, int64_t P, int prec) {
  MlpShape ms;
  if (make_shape(sd, &ms)) return -1;
  const int compact = prec == COPE_PREC_BF16 && sdf_infer_compact_bf16(ms);
  return cope_sdf_saved_floats(sd, P, compact ? 0 : 1, prec);
}

// forward-only workspace: on the tensor-core path the two forward entry points need the packed weights and the PE-gradient
// buffers only (7.5 kB per point with the saved blocks instead of 26.6 kB with the training workspace)
static int64_t infer_core_ws_floats(const cope_mlp_desc* sd, const cope_mlp_desc* cd, int64_t P, int prec) {
  if (prec != COPE_PREC_BF16) return cope_render_mlp_ws_floats(sd, cd, P, prec);
  MlpShape ms, mc;
  if (make_shape(sd, &ms) || make_shape(cd, &mc)) return -1;
  const int64_t a = sdf_fwd_ws_floats_bf16(ms, P), b = color_fwd_ws_floats_bf16(mc);
  if (a < 0 || b < 0) return -1;
  return (a > b ? a : b) + 1024;
}

int64_t cope_render_mlp_infer_ws_floats(const cope_mlp_desc* sd, const cope_mlp_desc* cd, int64_t P, int prec) {
  prec &= 0xFF;
  const int64_t w = infer_core_ws_floats(sd, cd, P, prec);
  const int64_t a = infer_sdf_saved_floats(sd, P, prec), b = cope_color_saved_floats(cd, P, prec);
  if (w < 0 || a < 0 || b < 0) return -1;
  return w + a + b + 256;
}

int cope_render_mlp_infer(const cope_mlp_desc* sd, const float* sdfW, const cope_mlp_desc* cd, const float* colW, const float* x,
                          const float* dirs, int dirs_group, int Lv, int64_t P, float* sdf, float* grad, float* rgb, float* ws,
                          int prec, cope_stream_t s_) {
  MlpShape ms, mc;
  if (make_shape(sd, &ms) || make_shape(cd, &mc)) return -1;
  if (P <= 0) return 0;
  const int flags = prec;
  prec &= 0xFF;
  const int64_t w = infer_core_ws_floats(sd, cd, P, prec);
  const int64_t a = infer_sdf_saved_floats(sd, P, prec);
  COPE_REQUIRE(w >= 0 && a >= 0, "render_mlp_infer: unsupported network shape");
  float* sdf_saved = ws + ((w + 63) / 64) * 64;
  float* col_saved = sdf_saved + ((a + 63) / 64) * 64;
  if (prec == COPE_PREC_BF16) {
    cudaStream_t s = as_stream(s_);
    int ld = 0;
    __nv_bfloat16* cin = color_cin_slot(mc, Lv, P, col_saved, &ld);
    COPE_REQUIRE(cin != nullptr, "render_mlp_infer: colour network shape not supported on the bf16 path");
    if (int rc = sdf_fwd_bf16(ms, sdfW, x, P, sdf, 1, nullptr, 0, grad, sdf_saved, ws, s, cin, ld, true, flat_pack_ptr(ms, sdfW, flags)))
      return rc;
    return color_fwd_bf16(mc, colW, x, dirs, dirs_group, Lv, grad, nullptr, 0, P, rgb, col_saved, ws, s, true, true,
                          flat_pack_ptr(mc, colW, flags));
  }
  return cope_render_mlp_fwd(sd, sdfW, cd, colW, x, dirs, dirs_group, Lv, P, sdf, grad, rgb, sdf_saved, col_saved, ws, prec, s_);
}

int cope_render_mlp_bwd(const cope_mlp_desc* sd, const float* sdfW, const cope_mlp_desc* cd, const float* colW, const float* x,
                        const float* dirs, int dirs_group, int Lv, int64_t P, const float* sdf_saved, const float* col_saved,
                        const float* d_sdf, float* d_grad, const float* d_rgb, float* dW_sdf, float* dW_col, float* dx,
                        float* ddirs_pp, float* ws, int prec, cope_stream_t s_) {
  MlpShape ms, mc;
  if (make_shape(sd, &ms) || make_shape(cd, &mc)) return -1;
  if (P <= 0) return 0;
  cudaStream_t s = as_stream(s_);
  const int flags = prec;
  prec &= 0xFF;
  const int64_t col_ws = cope_color_ws_floats(cd, P, prec);
  if (prec == COPE_PREC_BF16) {
    float* ws_sdf = ws + col_ws;
    int ld = 0;
    __nv_bfloat16* slot = sdf_bwd_dfeat_slot(ms, P, ws_sdf, &ld);
    COPE_REQUIRE(slot != nullptr, "render_mlp_bwd: SDF network shape not supported on the bf16 path");
    if (int rc = color_bwd_bf16(mc, colW, dirs, dirs_group, Lv, P, col_saved, d_rgb, dW_col, dx, ddirs_pp, d_grad, nullptr, 0, ws, s,
                                slot, ld, flat_pack_ptr(mc, colW, flags)))
      return rc;
    return sdf_bwd_bf16(ms, sdfW, x, P, sdf_saved, d_sdf, 1, nullptr, 0, d_grad, dW_sdf, dx, 1, ws_sdf, s, true,
                        flat_pack_ptr(ms, sdfW, flags));
  }
  const int F = ms.d_out - 1;
  float* dfeat = ws;
  float* rest = ws + P * (int64_t)F;
  if (int rc = cope_color_bwd(cd, colW, dirs, dirs_group, Lv, P, col_saved, d_rgb, dW_col, dx, ddirs_pp, d_grad, dfeat, F, rest, prec, s_))
    return rc;
  return cope_sdf_bwd(sd, sdfW, x, P, sdf_saved, d_sdf, 1, dfeat, F, d_grad, dW_sdf, dx, 1, rest, prec, s_);
}

}  // extern "C"
