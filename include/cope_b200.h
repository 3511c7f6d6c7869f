/*
 * cope_b200.h — C ABI of libcope_b200.so: the B200 (sm_100a) kernels behind the cope-nerf NeuS
 * render/train hot path.
 *
 * The reference (HoangChuongNguyen/cope-nerf) has NO native boundary: its hot path is eager PyTorch
 * inside model/neus_renderer.py, model/neus_fields.py, model/neus_embedder.py, model/poses_retriever.py,
 * model/common.py and model/training.py.  This ABI is therefore ours to define; each entry point names
 * the reference lines it replaces.  The host side (cope_nerf_b200/*.py) keeps the reference's nn.Module
 * API and calls these functions from torch.autograd.Function.forward/backward via ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; row-major, fp32, contiguous unless a stride is given
 *   - the caller allocates every buffer (including workspaces, sized by the *_floats queries)
 *   - the callee never allocates, frees or synchronises; all work is enqueued on `stream`
 *   - return 0 on success, negative on error; cope_last_error() gives the text (thread-local)
 *   - `cope_stream_t` is a cudaStream_t
 */
#ifndef COPE_B200_H
#define COPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cope_stream_t;

#define COPE_MAX_LIN 12

/* precision of the MLP paths */
#define COPE_PREC_FP32 0 /* fp32 SIMT GEMMs: strict-parity mode (<= 1e-3 rel vs the reference)          */
#define COPE_PREC_BF16 1 /* bf16 tcgen05 tensor-core path, fp32 accumulate (cos-sim > 0.999 contract)    */
/* OR-ed into `prec` of cope_sdf_query: the head of `ws` still holds the packed bf16 weights that an earlier cope_sdf_query left there
 * on the same stream for the same Wflat contents (the four sampling queries of one NeuSRenderer.forward): skip the re-pack launch */
#define COPE_WS_HOLDS_PACK 0x100
/* OR-ed into `prec` of any bf16 MLP entry point: the flat parameter buffer passed as Wflat is followed, at float offset
 * cope_mlp_pack_offset(desc), by the packed bf16 operands that cope_mlp_pack wrote there (cope_mlp_pack_floats floats).
 * The entry point then skips its own re-pack: one pack launch per network and step instead of five. */
#define COPE_FLAT_HAS_PACK 0x200

/* Shape of a weight-normalised MLP (SDFNetwork / RenderingNetwork, model/neus_fields.py:205-374).
 * Flat parameter layout (floats): for l in 0..n_lin-1: W_l [dims_out[l] x dims_in[l]] row-major, then
 * b_l [dims_out[l]].  Gradients use the same layout. */
typedef struct {
  int32_t n_lin;                 /* number of linear layers: 9 (SDF) or 5 (colour)                       */
  int32_t d_in;                  /* raw input width before PE: 4 = (x,y,z,t)                              */
  int32_t multires;              /* PE frequencies L (6 for SDF); PE width = d_in*(1+2L)                  */
  int32_t skip_layer;            /* layer whose input is cat([h, pe])/sqrt(2) (4 for SDF), -1 = none     */
  int32_t dims_in[COPE_MAX_LIN]; /* K of each layer                                                      */
  int32_t dims_out[COPE_MAX_LIN];/* N of each layer (layer skip_layer-1 emits dims_in[skip]-pe_width)    */
  int32_t activation;            /* hidden activation of the cope_sdf_* family: COPE_ACT_SOFTPLUS100 (SDFNetwork) or
                                    COPE_ACT_LEAKY_RELU (MotionNetwork, model/neus_fields.py:140); strict fp32 path only */
  float act_param;               /* negative slope of COPE_ACT_LEAKY_RELU (0.2)                           */
} cope_mlp_desc;

#define COPE_ACT_SOFTPLUS100 0
#define COPE_ACT_LEAKY_RELU 1

int cope_version(void);
const char* cope_last_error(void);
/* kernels launched by this library since load (monotonic; bench.py reports the per-run delta) */
uint64_t cope_launch_count(void);
/* number of floats in the flat [W_l | b_l]* buffer of an MLP */
int64_t cope_mlp_flat_floats(const cope_mlp_desc* d);

/* ---- weight norm: W = g * v / ||v||_row  (nn.utils.weight_norm, model/neus_fields.py:261-262,339-340) */
int cope_weightnorm_fwd(const float* v, const float* g, float* W, int rows, int cols, cope_stream_t s);
/* dv, dg are OVERWRITTEN */
int cope_weightnorm_bwd(const float* v, const float* g, const float* dW, float* dv, float* dg, int rows,
                        int cols, cope_stream_t s);

/* All n layers of one network in one launch (host arrays of device pointers / sizes): flat = [W_l | b_l]* with
 * W_l = g_l * v_l / ||v_l||_row at float offset w_off[l] and the bias copied to b_off[l]; the backward writes
 * dv_l, dg_l, db_l from dflat: OVERWRITTEN, or with accumulate != 0 ADDED to (the buffers then are the parameters' gradient
 * buffers themselves, e.g. views of one flat all-reduce bucket).  n <= 16. */
int cope_flat_weights_fwd(int n, const void* const* v, const void* const* g, const void* const* b, const int* rows,
                          const int* cols, const int64_t* w_off, const int64_t* b_off, float* flat, cope_stream_t s);
int cope_flat_weights_bwd(int n, const void* const* v, const void* const* g, const int* rows, const int* cols,
                          const int64_t* w_off, const int64_t* b_off, const float* dflat, void* const* dv, void* const* dg,
                          void* const* db, int accumulate, cope_stream_t s);

/* ---- positional encoding (model/neus_embedder.py:6-51): out [P x d*(1+2L)] */
int cope_embed_fwd(const float* x, int64_t P, int d, int L, float* out, cope_stream_t s);

/* ---- SDF network (model/neus_fields.py:268-303) --------------------------------------------------- */
/* floats of `saved` needed by cope_sdf_fwd (with_grad: also keeps the reverse-sweep deltas) */
int64_t cope_sdf_saved_floats(const cope_mlp_desc* d, int64_t P, int with_grad, int prec);
/* floats of scratch workspace needed by query / fwd / bwd */
int64_t cope_sdf_ws_floats(const cope_mlp_desc* d, int64_t P, int prec);

/* floats of workspace needed by cope_sdf_query alone (much smaller than the training workspace) */
int64_t cope_sdf_query_ws_floats(const cope_mlp_desc* d, int64_t P, int prec);
/* sdf only, no state kept: SDFNetwork.sdf under no_grad (neus_renderer.py:499, :292). sdf_out [P] */
int cope_sdf_query(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, float* sdf_out,
                   float* ws, int prec, cope_stream_t s);
/* SDFNetwork.forward(x): column 0 -> sdf (row stride sdf_ld), columns 1.. -> feat (row stride feat_ld); for one
 * [P x d_out] tensor y pass (y, d_out, y+1, d_out).  grad [P x d_in] = SDFNetwork.gradient(x) (analytic reverse
 * sweep, replaces the autograd.grad of neus_fields.py:291-303) or NULL.  `saved` keeps activations for bwd. */
int cope_sdf_fwd(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, float* sdf, int sdf_ld,
                 float* feat, int feat_ld, float* grad, float* saved, float* ws, int prec, cope_stream_t s);
/* Backward of both outputs.  Upstream of forward(): d_sdf (row stride d_sdf_ld) and d_feat (row stride d_feat_ld),
 * either may be NULL (= zero); dgrad [P x d_in] or NULL (second-order path: the double backward that the eikonal
 * loss / colour-net normals need).  dWflat is ACCUMULATED into.
 * dx [P x d_in] or NULL receives (ACCUMULATED if dx_accumulate) the gradient w.r.t. x THROUGH THE VALUE PATH ONLY,
 * matching render_core where .gradient() sees pts_time.detach() (neus_renderer.py:352-356). */
int cope_sdf_bwd(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, const float* saved,
                 const float* d_sdf, int d_sdf_ld, const float* d_feat, int d_feat_ld, const float* dgrad,
                 float* dWflat, float* dx, int dx_accumulate, float* ws, int prec, cope_stream_t s);

/* ---- colour network (model/neus_fields.py:346-374, mode 'idr') ------------------------------------ */
int64_t cope_color_saved_floats(const cope_mlp_desc* d, int64_t P, int prec);
int64_t cope_color_ws_floats(const cope_mlp_desc* d, int64_t P, int prec);
/* input = cat[x(4) | PE_Lv(dirs)(3+6Lv) | normals(4) | feat(d_feat)].  dirs is [P/dirs_group x 3]: one
 * direction per `dirs_group` consecutive points (= samples per ray; 1 for per-point dirs).
 * feat has row stride feat_ld (257 when it aliases y[:,1:]).  rgb [P x 3] (sigmoid applied). */
int cope_color_fwd(const cope_mlp_desc* d, const float* Wflat, const float* x, const float* dirs,
                   int dirs_group, int Lv, const float* normals, const float* feat, int feat_ld, int64_t P,
                   float* rgb, float* saved, float* ws, int prec, cope_stream_t s);
/* d_rgb [P x 3].  Outputs (any may be NULL): dx [P x 4] ACCUMULATED, ddirs [P x 3] per point OVERWRITTEN,
 * dnormals [P x 4] ACCUMULATED, dfeat (row stride dfeat_ld) OVERWRITTEN.  dWflat ACCUMULATED. */
int cope_color_bwd(const cope_mlp_desc* d, const float* Wflat, const float* dirs, int dirs_group, int Lv,
                   int64_t P, const float* saved, const float* d_rgb, float* dWflat, float* dx, float* ddirs,
                   float* dnormals, float* dfeat, int dfeat_ld, float* ws, int prec, cope_stream_t s);

/* ---- render_core's MLP stage in one call (model/neus_renderer.py:352-358): sdf [P], grad [P x 4] (= .gradient()),
 * rgb [P x 3].  On the bf16 path the feature vector / its gradient are handed between the two networks in bf16.
 * Backward: d_sdf [P] (or NULL), d_grad [P x 4] holds the upstream gradient of `grad` on entry and is ACCUMULATED with
 * the colour net's contribution, d_rgb [P x 3]; dW_sdf / dW_col ACCUMULATED; dx [P x 4] ACCUMULATED (value path only) or
 * NULL; ddirs_pp [P x 3] per-point view-direction gradient (OVERWRITTEN) or NULL. */
int64_t cope_render_mlp_ws_floats(const cope_mlp_desc* sdf_desc, const cope_mlp_desc* col_desc, int64_t P, int prec);
int cope_render_mlp_fwd(const cope_mlp_desc* sdf_desc, const float* sdfW, const cope_mlp_desc* col_desc, const float* colW,
                        const float* x, const float* dirs, int dirs_group, int Lv, int64_t P, float* sdf, float* grad,
                        float* rgb, float* sdf_saved, float* col_saved, float* ws, int prec, cope_stream_t s);
int cope_render_mlp_bwd(const cope_mlp_desc* sdf_desc, const float* sdfW, const cope_mlp_desc* col_desc, const float* colW,
                        const float* x, const float* dirs, int dirs_group, int Lv, int64_t P, const float* sdf_saved,
                        const float* col_saved, const float* d_sdf, float* d_grad, const float* d_rgb, float* dW_sdf,
                        float* dW_col, float* dx, float* ddirs_pp, float* ws, int prec, cope_stream_t s);

/* Inference form of cope_render_mlp_fwd (full-image rendering: model/training.py:210-283, eval.py:133-157): same sdf / grad /
 * rgb, nothing is kept for a backward pass (on the bf16 path the reverse-sweep deltas, the colour net's hidden activations
 * and its input tail are never written to HBM).  `ws` holds cope_render_mlp_infer_ws_floats floats. */
int64_t cope_render_mlp_infer_ws_floats(const cope_mlp_desc* sdf_desc, const cope_mlp_desc* col_desc, int64_t P, int prec);
int cope_render_mlp_infer(const cope_mlp_desc* sdf_desc, const float* sdfW, const cope_mlp_desc* col_desc, const float* colW,
                          const float* x, const float* dirs, int dirs_group, int Lv, int64_t P, float* sdf, float* grad,
                          float* rgb, float* ws, int prec, cope_stream_t s);

/* test hook (bf16 path): float offset, inside the `ws` of cope_render_mlp_bwd, of the two [P x 64] fp32 buffers that hold
 * the value-path gradient w.r.t. the positional encoding (layer 0 / skip layer) before the PE backward folds them into dx */
int64_t cope_dbg_render_bwd_eb_offset(const cope_mlp_desc* sdf_desc, const cope_mlp_desc* col_desc, int64_t P);

/* ---- ray points (neus_renderer.py:337-350 / :495-498 / :285) --------------------------------------
 * pts_time [N*S x 4] = (o + d * zz, t) with zz = z + dists/2 if use_mid else z.
 * dists / mid_z [N x S] may be NULL.  The last interval is sample_dist = (far[0]-near[0])/n_coarse,
 * read on the device (near/far are device pointers; no host sync). */
int cope_ray_points(const float* rays_o, const float* rays_d, const float* z, const float* time_step,
                    const float* near, const float* far, int n_coarse, int64_t N, int S, int use_mid,
                    float* pts_time, float* dists, float* mid_z, cope_stream_t s);
/* d_pts [N*S x 4] -> d_rays_o [N x 3] (overwritten), d_rays_d [N x 3] (ACCUMULATED: compositing adds the
 * true_cos term first).  d_dirs_pp [N*S x 3] or NULL: per-point view-direction gradients (colour net) summed
 * per ray into d_rays_d. */
int cope_ray_points_bwd(const float* d_pts, const float* mid_z, const float* d_dirs_pp, int64_t N, int S,
                        float* d_rays_o, float* d_rays_d, cope_stream_t s);
/* coarse stratified depths (neus_renderer.py:466-483): t_rand [N x S] or NULL (eval) */
int cope_coarse_z(const float* near, const float* far, const float* t_rand, int64_t N, int S, float* z,
                  cope_stream_t s);

/* ---- hierarchical sampling ----------------------------------------------------------------------- */
/* sample_pdf's inverse-CDF step given a CDF (neus_renderer.py:47-70, det=True): bit-exact contract.
 * cdf, bins [N x S]; out samples [N x K], inds int64 [N x K] (searchsorted right=True) or NULL */
int cope_sample_cdf(const float* cdf, const float* bins, int64_t N, int S, int K, float* samples,
                    int64_t* inds, cope_stream_t s);
/* up_sample (neus_renderer.py:178-224) incl. sample_pdf: z, sdf [N x S] -> new_z [N x K].
 * Optional debug outputs: cdf [N x S], inds [N x K]. */
int cope_upsample(const float* z, const float* sdf, int64_t N, int S, int K, float inv_s, float* new_z,
                  float* cdf_out, int64_t* inds_out, cope_stream_t s);
/* cat_z_vals without the MLP query (neus_renderer.py:286-297): merge sorted z [N x S] with new_z [N x K];
 * sdf/new_sdf may be NULL (last step).  Ties: old before new. */
int cope_merge_z(const float* z, const float* new_z, const float* sdf, const float* new_sdf, int64_t N, int S,
                 int K, float* z_out, float* sdf_out, cope_stream_t s);

/* ---- compositing (neus_renderer.py:360-420, :565-575) --------------------------------------------- */
/* Inputs per sample: sdf [P], grad [P x 4] (normal = first 3), rgb [P x 3], z/dists [N x S]; per ray: rays_d
 * [N x 3], rays_d_norm [N] (used only if eval_mode).  variance: device scalar (SingleVarianceNetwork).
 * Outputs: weights [N x S], color [N x 3], depth [N] (divided by |d| if eval), weighted_z [N] (never divided),
 * cdf [N x S] (prev_cdf, logging), wsum [N], wmax [N],
 * inv_s_out [1]. */
int cope_composite_fwd(const float* sdf, const float* grad, const float* rgb, const float* z,
                       const float* dists, const float* rays_d, const float* rays_d_norm,
                       const float* variance, float cos_anneal, int eval_mode, int64_t N, int S,
                       float* weights, float* color, float* depth, float* weighted_z, float* cdf, float* wsum,
                       float* wmax, float* inv_s_out, cope_stream_t s);
/* Upstream: d_color [N x 3], d_depth [N], d_weights [N x S], d_grad_in [P x 4] = upstream gradients of
 * `normals` / `sdf_flows` (any may be NULL = zero; d_grad_in may alias d_grad).
 * Outputs: d_sdf [P] OVERWRITTEN, d_grad [P x 4] OVERWRITTEN (= d_grad_in + the true_cos term: no pre-fill, no
 * read-modify-write when there is no upstream gradient), d_rgb [P x 3] OVERWRITTEN, d_variance [1] ACCUMULATED,
 * d_rays_d [N x 3] OVERWRITTEN (the true_cos term). */
int cope_composite_bwd(const float* sdf, const float* grad, const float* rgb, const float* z,
                       const float* dists, const float* rays_d, const float* rays_d_norm,
                       const float* variance, float cos_anneal, int eval_mode, int64_t N, int S,
                       const float* d_color, const float* d_depth,
                       const float* d_weights, const float* d_grad_in, float* d_sdf, float* d_grad, float* d_rgb,
                       float* d_variance, float* d_rays_d, cope_stream_t s);

/* ---- pose + ray generation (poses_retriever.py:25-32, common.py:175-215,255-308, training.py:474-487) */
/* c2w [4x4] = make_c2w(r, t) @ init_c2w   (r, t, init_c2w: device pointers to one camera's entries) */
int cope_pose_fwd(const float* r, const float* t, const float* init_c2w, float* c2w, cope_stream_t s);
int cope_pose_bwd(const float* r, const float* t, const float* init_c2w, const float* d_c2w, float* dr,
                  float* dt, cope_stream_t s);
/* rays from normalised pixels [N x 2]: M = inv(scale) inv(world) inv(camera); o = M[:,3], d = normalise(M p - o)
 * outputs rays_o [N x 3], rays_d [N x 3], rays_d_norm [N] */
int cope_raygen_fwd(const float* pixels, const float* camera_mat, const float* world_mat,
                    const float* scale_mat, int64_t N, float* rays_o, float* rays_d, float* rays_d_norm,
                    cope_stream_t s);
/* d_world [4x4] OVERWRITTEN with dL/d world_mat given d_rays_o, d_rays_d (d_norm may be NULL) */
int cope_raygen_bwd(const float* pixels, const float* camera_mat, const float* world_mat,
                    const float* scale_mat, int64_t N, const float* d_rays_o, const float* d_rays_d,
                    const float* d_norm, float* d_world, float* ws /* >= 16 floats, zeroed by callee */,
                    cope_stream_t s);

/* ---- generic fp32 GEMM (used by the tests to exercise the kernel directly) ------------------------
 * C[M x N] (+)= op(A) op(B); transA: A stored [K x M]; transB: B stored [N x K] */
int cope_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, int accumulate, cope_stream_t s);

/* ---- bf16 tcgen05 building blocks (exposed for kernel-level tests; the MLP entry points above use them when
 * prec == COPE_PREC_BF16).  Packed weight layout: [Kp/8][Np][8] bf16 = UMMA no-swizzle K-major core matrices.
 * Packed row n < n_src / packed k < k_src come from W (row stride ldw), the rest is zero padding; transposed packs W^T.
 * epi: 0 store, 1 bias+softplus(beta=100), 2 bias+relu, 3 bias+sigmoid. */
int cope_tc_pack(const float* W, int ldw, int n_src, int k_src, int Np, int Kp, int transposed, void* out_bf16,
                 cope_stream_t s);
/* out[M x N] = epi(A[M x K] (bf16, row stride lda) * Wpacked^T + bias); N mult of 16 <= 256, K mult of 64 <= 320 */
int cope_tc_gemm(int M, int N, int K, const void* A_bf16, int lda, const void* Bp_bf16, const float* bias, int epi,
                 float alpha, void* out, int ldo, int out_f32, cope_stream_t s);
/* dW[m_valid x n_valid] (fp32, row stride ldw) += X[P x Mp]^T Y[P x Np] (bf16); Mp in {128,256}, Np mult of 16.
 * ws: cope_tc_wgrad_ws_floats() floats of scratch for the per-CTA partial tiles (summed without atomics). */
int64_t cope_tc_wgrad_ws_floats(void);
int cope_tc_wgrad(int64_t P, int Mp, int Np, int m_valid, int n_valid, const void* X, int ldx, const void* Y, int ldy,
                  float* dW, int ldw, float* ws, cope_stream_t s);

/* ---- continuous pose model (MotionNetwork, model/neus_fields.py:142-183) -------------------------
 * The network itself is a cope_sdf_fwd / cope_sdf_bwd MLP with activation = COPE_ACT_LEAKY_RELU (strict fp32 path).
 * cope_pose_integrate_*: wv [F * n_sub x 6] = (angular velocity, velocity) at the n_sub time samples of each of the F
 * consecutive frame pairs, dt [F] = the time interval of each pair; rel [F x 16] row-major 4x4 relative poses
 * (compute_consecutive_relative_pose, :142-160).  Backward: d_wv and d_dt [F] (may be NULL) OVERWRITTEN.
 * cope_pose_chain_*: w2c [(F+1) x 16], w2c_0 = I, w2c_{i+1} = rel_i @ w2c_i (compute_w2c_mappings, :172-183);
 * backward: d_w2c [(F+1) x 16] -> d_rel [F x 16] OVERWRITTEN. */
int cope_pose_integrate_fwd(const float* wv, const float* dt, int F, int n_sub, float* rel, cope_stream_t s);
int cope_pose_integrate_bwd(const float* wv, const float* dt, int F, int n_sub, const float* d_rel, float* d_wv, float* d_dt,
                            cope_stream_t s);
int cope_pose_chain_fwd(const float* rel, int F, float* w2c, cope_stream_t s);
int cope_pose_chain_bwd(const float* rel, const float* w2c, int F, const float* d_w2c, float* d_rel, cope_stream_t s);

/* ---- per-ray image reductions of the evaluation render (model/training.py:236-283): normal[n] = R * sum_s w[n,s] * grad[n,s,:3]
 * and depth_hw[n] = -(world_mat @ [pts[n, argmax_s w], 1]).z, world_mat = device pointer to a row-major 4x4 (R = its 3x3).
 * Optional (flow_affine != NULL): the predicted forward optical flow of :265-283, flow[n] = (proj(KS * F [sum_s w p; sum_s w]) -
 * pix_norm[n]) * (flow_sx, flow_sy), where F [3 x 4] is the scene-flow integration of the sub-steps composed into one affine map
 * (each Euler sub-step p <- p + dt (w x p + v) is affine in p), KS [3 x 3] = scale_mat[:3,:3] @ camera_mat[:3,:3], pix_norm
 * [N x 2] the rays' normalised pixels and flow_sx / flow_sy = w / 2, h / 2 (:296-297). */
int cope_eval_reduce(const float* weights, const float* grad, const float* pts, const float* world_mat, int64_t N, int S,
                     float* normal_out, float* depth_hw_out, const float* flow_affine, const float* KS, const float* pix_norm,
                     float flow_sx, float flow_sy, float* flow_out, cope_stream_t s);

/* ---- loss reductions of the training step (SURVEY.md 8 a17 + 8f rank 2), one forward + one backward launch per group -------
 * cope_step_losses_*: rgb L1 (model/training.py:508), eikonal (train.py:526) and, when `motion` != NULL, the SDF-flow loss
 * (train.py:467-477).  color / rgb_gt [N x 3]; grad4 [P x 4] = (d sdf/dx, d sdf/dy, d sdf/dz, d sdf/dt); pts4 [P x 4];
 * weights [P] (treated as constants, as the reference detaches them); motion [6] = angular velocity | velocity (device);
 * w_sum_global [1] or NULL = normaliser of the SDF-flow term when rays are sharded over ranks.
 * losses [4] = (w_rgb l_rgb + w_eik l_eik + w_flow l_flow, l_rgb, l_eik, l_flow); coef [4] = the normalised loss weights the
 * backward multiplies with (and sum w); ws >= 8 floats, zeroed by the callee.
 * Backward: g [1] = dL/d losses[0] (device; NULL = 1); d_color [N x 3], d_grad4 [P x 4], d_pts4 [P x 4] (may be NULL)
 * OVERWRITTEN; d_motion [6] ACCUMULATED (may be NULL). */
int cope_step_losses_fwd(const float* color, const float* rgb_gt, const float* grad4, const float* pts4, const float* weights,
                         const float* motion, const float* w_sum_global, int64_t N, int64_t P, float w_rgb, float w_eik,
                         float w_flow, float* losses, float* coef, float* ws, cope_stream_t s);
int cope_step_losses_bwd(const float* color, const float* rgb_gt, const float* grad4, const float* pts4, const float* weights,
                         const float* motion, int64_t N, int64_t P, const float* coef, const float* g, float* d_color,
                         float* d_grad4, float* d_pts4, float* d_motion, cope_stream_t s);
/* wp [N x 4] = sum_s weights[n,s] * (pts4[n,s].xyz, 1): the per-sample part of the flow-RGB loss (train.py:488-489).
 * Backward: d_weights [N x S] and d_pts4 [P x 4] OVERWRITTEN (either may be NULL). */
int cope_weighted_points_fwd(const float* weights, const float* pts4, int64_t N, int S, float* wp, cope_stream_t s);
int cope_weighted_points_bwd(const float* weights, const float* pts4, const float* d_wp, int64_t N, int S, float* d_weights,
                             float* d_pts4, cope_stream_t s);
/* Flow-RGB loss (train.py:486-517 + warp_pixel :235-244): for each of the T reference frames, m = w2c_t[:3,:] wp,
 * q = KS_t m, pixel flow = ((q.xy / q.z) - norm_pix) * (W/2, H/2), source = pix + flow, bilinear border-clamped sample of
 * ref_imgs [T x 3 x H x W] (grid_sample, align_corners=True), loss = sum_t (sum_valid |warped - rgb_gt| / (#valid + 1e-10)) / 3.
 * flow_pred [T x N x 2] may be NULL; ws >= 2T + 1 floats, zeroed by the callee and read again by the backward.
 * Backward: g [1] device (NULL = 1); d_wp [N x 4] OVERWRITTEN; d_w2c [T x 16] ACCUMULATED (may be NULL). */
int cope_flow_rgb_fwd(const float* wp, const float* w2c, const float* KS, const float* norm_pix, const float* pix,
                      const float* ref_imgs, const float* rgb_gt, int64_t N, int T, int H, int W, float* flow_pred, float* loss,
                      float* ws, cope_stream_t s);
int cope_flow_rgb_bwd(const float* wp, const float* w2c, const float* KS, const float* norm_pix, const float* pix,
                      const float* ref_imgs, const float* rgb_gt, int64_t N, int T, int H, int W, const float* ws, const float* g,
                      float* d_wp, float* d_w2c, cope_stream_t s);

/* ---- weights packed once per step (bf16 path; see COPE_FLAT_HAS_PACK).  is_color: 0 = SDF network layout (forward + transposed
 * blocks + the sdf row), 1 = colour network layout (Lv = multires_view).  flat_with_tail: the flat fp32 parameters [W_l | b_l]*
 * with cope_mlp_pack_floats extra floats allocated from offset cope_mlp_pack_offset on. */
int64_t cope_mlp_pack_offset(const cope_mlp_desc* d);
int64_t cope_mlp_pack_floats(const cope_mlp_desc* d, int is_color, int Lv);
int cope_mlp_pack(const cope_mlp_desc* d, int is_color, int Lv, float* flat_with_tail, cope_stream_t s);

/* ---- depth-patch smoothness (model/losses.py:7-18 SmoothnessLoss, :20-38 EdgePreservingSmoothnessLoss; called on
 * depth_pred.view(-1, ps, ps, 1) and rgb_gt.view(-1, ps, ps, 3) at train.py:519-525) ------------------------------------
 * depth [n_patches x ps x ps], rgb [n_patches x ps x ps x 3] (NULL: no edge-aware term), 2 <= ps <= 8.
 * losses[3] = { w_edge * edge + w_smooth * smooth, edge, smooth }; ws: 12 floats of scratch.  The backward OVERWRITES
 * d_depth [n_patches x ps x ps] with g[0] * d losses[0] / d depth (g: device scalar, NULL = 1). */
int cope_patch_smooth_fwd(const float* depth, const float* rgb, int64_t n_patches, int ps, float gamma, float w_edge,
                          float w_smooth, float* losses, float* ws, cope_stream_t s);
int cope_patch_smooth_bwd(const float* depth, const float* rgb, int64_t n_patches, int ps, float gamma, float w_edge,
                          float w_smooth, const float* g, float* d_depth, cope_stream_t s);

/* ---- photometric relative-pose refinement (utils_poses/pose_refinement.py:34-61 compute_loss_and_warp_image) --------------
 * images / next_images [B x 3 x H x W], depths [B x H x W] (of `images`), K [B x 3 x 3] (maps camera xyz to the normalised
 * [-1,1] pixel grid of pose_refinement.py:88-96), poses [B x 4 x 4] relative camera poses.  loss[0] = sum |warp(next_images) -
 * images| valid / sum valid; warped [B x 3 x H x W] may be NULL; ws: 4 floats, kept for the backward.  The backward ACCUMULATES
 * g[0] * d loss / d poses into d_poses [B x 4 x 4] (rows 0..2; the caller zero-fills). */
int cope_pose_refine_fwd(const float* images, const float* next_images, const float* depths, const float* K, const float* poses,
                         int B, int H, int W, float* warped, float* loss, float* ws, cope_stream_t s);
int cope_pose_refine_bwd(const float* images, const float* next_images, const float* depths, const float* K, const float* poses,
                         int B, int H, int W, const float* ws, const float* g, float* d_poses, cope_stream_t s);

/* ---- training-pixel selection (process_data, model/training.py:413-471; arange_pixels, model/common.py:12-39) ----------
 * n_patches patch_size x patch_size patches of an h x w frame -> N = n_patches * patch_size^2 rays, row-major inside each patch:
 * ray_idx [N] int64 flat pixel ids, pix [N x 2] (col, row) as floats, norm_pix [N x 2] = 2 * (col, row) / (w-1, h-1) - 1,
 * rgb_gt [N x 3] gathered from img [3 x h x w] (any output may be NULL).  corners [n_patches] int64 = top-left corner ids
 * in the (h-ps+1) x (w-ps+1) grid (the reference's randperm prefix); corners == NULL draws distinct corners on the device from
 * a keyed permutation of that grid (seed). */
int cope_sample_pixels(const int64_t* corners, uint64_t seed, int h, int w, int patch_size, int n_patches, const float* img,
                       int64_t* ray_idx, float* pix, float* norm_pix, float* rgb_gt, cope_stream_t s);

/* ---- optimiser step (train.py:59-60 builds torch.optim.Adam, model/training.py:552-558 steps it) ------------------------
 * torch.optim.Adam (no amsgrad) over ONE flat fp32 buffer of n elements: param, grad, exp_avg, exp_avg_sq [n].
 *   exp_avg += (1 - beta1) (grad - exp_avg);  exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) grad^2;
 *   param -= lr / (1 - beta1^t) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps)
 * step [1] is the device-resident step count t >= 1 (the caller increments it before the call, so CUDA-graph replays advance
 * it); weight_decay != 0 adds weight_decay * param to the gradient first (L2 penalty, as torch.optim.Adam). */
int cope_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* step, float lr,
                   float beta1, float beta2, float eps, float weight_decay, cope_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* COPE_B200_H */
