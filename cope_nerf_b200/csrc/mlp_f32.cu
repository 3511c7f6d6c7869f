// SDF / colour MLPs, strict-parity fp32 path: positional-encoding kernels, the analytic reverse sweep that
// replaces autograd.grad (neus_fields.py:291-303), and the first + second order backward.
//
// Notation (per point):  in_0 = PE(x);  z_l = W_l in_l + b_l;  in_{l+1} = softplus(z_l)   (skip layer:
// in_s = [softplus(z_{s-1}) | PE] / sqrt2).  Reverse sweep: delta_l = a_{l+1} * softplus'(z_l),
// a_l = W_l^T delta_l, grad = J_PE^T (a_0 + a_s[pe part]/sqrt2).
// Backward of `grad` (upstream G) is a forward-mode tangent pass: t_0 = J_PE G, u_l = W_l t_l,
// t_{l+1} = u_l * softplus'(z_l), which contributes  zb2_l = u_l * delta_l * 100 (1 - softplus'(z_l))  to the
// adjoint of z_l and  dW_l += delta_l (x) t_l.
#include <algorithm>

#include "common.cuh"
#include "mlp_shape.cuh"

namespace cope {

// ------------------------------------------------------------------------------------------- PE kernels
// out[p, :] = [x | sin(2^k x) | cos(2^k x)]_k ; optional second destination scaled by 1/sqrt2 (skip concat)
__global__ void pe_fwd_kernel(const float* __restrict__ x, int64_t P, int d, int L, int x_group,
                              float* __restrict__ out, int ld, float* __restrict__ out2, int ld2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d) return;
  int64_t p = i / d;
  int dd = (int)(i - p * d);
  float v = x[(p / x_group) * d + dd];
  float* o = out + p * ld;
  o[dd] = v;
  float* o2 = out2 ? out2 + p * ld2 : nullptr;
  if (o2) o2[dd] = v * kInvSqrt2;
  float f = 1.0f;
  for (int k = 0; k < L; ++k, f *= 2.0f) {
    float s, c;
    sincosf(v * f, &s, &c);
    o[d * (1 + 2 * k) + dd] = s;
    o[d * (2 + 2 * k) + dd] = c;
    if (o2) {
      o2[d * (1 + 2 * k) + dd] = s * kInvSqrt2;
      o2[d * (2 + 2 * k) + dd] = c * kInvSqrt2;
    }
  }
}

// g[p, dd] (+)= J_PE^T (ge0 + ge1)[p, :]
__global__ void pe_vjp_kernel(const float* __restrict__ x, int64_t P, int d, int L, const float* __restrict__ ge0,
                              int ld0, const float* __restrict__ ge1, int ld1, float* __restrict__ g, int ldg,
                              int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d) return;
  int64_t p = i / d;
  int dd = (int)(i - p * d);
  float v = x[p * d + dd];
  const float* a = ge0 + p * ld0;
  const float* b = ge1 ? ge1 + p * ld1 : nullptr;
  auto at = [&](int c) { return a[c] + (b ? b[c] : 0.0f); };
  float acc = at(dd);
  float f = 1.0f;
  for (int k = 0; k < L; ++k, f *= 2.0f) {
    float s, c;
    sincosf(v * f, &s, &c);
    acc += f * (c * at(d * (1 + 2 * k) + dd) - s * at(d * (2 + 2 * k) + dd));
  }
  float* o = g + p * ldg + dd;
  *o = accumulate ? *o + acc : acc;
}

// t0[p, :] = J_PE G[p, :]; optional copy scaled by 1/sqrt2 (tangent of the skip concat)
__global__ void pe_jvp_kernel(const float* __restrict__ x, int64_t P, int d, int L, const float* __restrict__ G,
                              float* __restrict__ t0, int ld, float* __restrict__ t2, int ld2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d) return;
  int64_t p = i / d;
  int dd = (int)(i - p * d);
  float v = x[p * d + dd], gv = G[p * d + dd];
  float* o = t0 + p * ld;
  float* o2 = t2 ? t2 + p * ld2 : nullptr;
  o[dd] = gv;
  if (o2) o2[dd] = gv * kInvSqrt2;
  float f = 1.0f;
  for (int k = 0; k < L; ++k, f *= 2.0f) {
    float s, c;
    sincosf(v * f, &s, &c);
    float ts = f * c * gv, tc = -f * s * gv;
    o[d * (1 + 2 * k) + dd] = ts;
    o[d * (2 + 2 * k) + dd] = tc;
    if (o2) {
      o2[d * (1 + 2 * k) + dd] = ts * kInvSqrt2;
      o2[d * (2 + 2 * k) + dd] = tc * kInvSqrt2;
    }
  }
}

// D[p, n] = w[n] * softplus'(Z[p, n])   (top of the reverse sweep: a_last = row 0 of the last W)
__global__ void bcast_sigp_kernel(const float* __restrict__ w, const float* __restrict__ Z, int ldz,
                                  float* __restrict__ D, int ldd, int64_t P, int n, int act, float slope) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * n) return;
  int64_t p = i / n;
  int c = (int)(i - p * n);
  D[p * ldd + c] = w[c] * act_d1(act, slope, Z[p * ldz + c]);
}

// out[c] += sum_p X[p, c]
__global__ void colsum_atomic_kernel(const float* __restrict__ X, int ld, int64_t P, int n, int rows_per_block,
                                     float* __restrict__ out) {
  int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    float acc = 0.0f;
    for (int64_t p = p0; p < p1; ++p) acc += X[p * ld + c];
    atomicAdd(out + c, acc);
  }
}

__global__ void strided_copy_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                    int64_t P, int n, float scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * n) return;
  int64_t p = i / n;
  int c = (int)(i - p * n);
  dst[p * ldd + c] = src[p * lds + c] * scale;
}

// dyb[p, :] = [d_sdf[p] | d_feat[p, :]]  (either may be null = zero)
__global__ void pack_dy_kernel(const float* __restrict__ d_sdf, int ld_s, const float* __restrict__ d_feat, int ld_f,
                               int64_t P, int d_out, float* __restrict__ dyb, int ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d_out) return;
  int64_t p = i / d_out;
  int c = (int)(i - p * d_out);
  float v = c == 0 ? (d_sdf ? d_sdf[p * ld_s] : 0.0f) : (d_feat ? d_feat[p * ld_f + (c - 1)] : 0.0f);
  dyb[p * ld + c] = v;
}

static inline dim3 grid1d(int64_t n, int bs = 256) { return dim3((unsigned)ceil_div(n, bs)); }

static int colsum(const float* X, int ld, int64_t P, int n, float* out, cudaStream_t s) {
  if (P <= 0 || n <= 0) return 0;
  const int rpb = 256;
  colsum_atomic_kernel<<<grid1d(P, rpb), 256, 0, s>>>(X, ld, P, n, rpb, out);
  COPE_CHECK_LAUNCH("colsum");
  return 0;
}

// dW[out x in] += Zb^T [out x P] * In [P x in]   (split-K over points, atomics)
static int wgrad(const float* Zb, int ldz, const float* In, int ldi, int64_t P, int out, int in, float* dW,
                 cudaStream_t s) {
  GemmArgs g = gemm_args(out, in, (int)P, Zb, ldz, In, ldi, dW, in);
  g.epi = EPI_ATOMIC;
  int tiles = (int)(ceil_div(out, 128) * ceil_div(in, 128));
  g.split_k = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(P, 256), ceil_div(4 * 148, tiles)));
  return launch_gemm(true, false, g, s);
}

struct SdfSaved {   // views into the caller's `saved` buffer
  float* pe; float* Z; float* H; float* D; int64_t P; int ldh, pe_w;
  float* z(int l) const { return Z + (int64_t)l * P * ldh; }        // l = 0..n_lin-2
  float* h(int l) const { return H + (int64_t)(l - 1) * P * ldh; }  // input of layer l, l = 1..n_lin-1
  float* dl(int l) const { return D + (int64_t)l * P * ldh; }       // delta_l, l = 0..n_lin-2
  const float* in(int l) const { return l == 0 ? pe : h(l); }
  int ld_in(int l) const { return l == 0 ? pe_w : ldh; }
};
static SdfSaved sdf_saved(const MlpShape& m, int64_t P, float* base) {
  SdfSaved v; v.P = P; v.ldh = m.ldh; v.pe_w = m.pe_w;
  int64_t hid = (int64_t)(m.n_lin - 1) * P * m.ldh;
  v.pe = base; v.Z = base + P * m.pe_w; v.H = v.Z + hid; v.D = v.H + hid;
  return v;
}

}  // namespace cope

using namespace cope;

extern "C" {

int64_t cope_mlp_flat_floats(const cope_mlp_desc* d) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  return m.n_flat;
}

int cope_embed_fwd(const float* x, int64_t P, int d, int L, float* out, cope_stream_t s) {
  if (P <= 0) return 0;
  pe_fwd_kernel<<<grid1d(P * d), 256, 0, as_stream(s)>>>(x, P, d, L, 1, out, d * (1 + 2 * L), nullptr, 0);
  COPE_CHECK_LAUNCH("pe_fwd");
  return 0;
}

int64_t cope_sdf_saved_floats(const cope_mlp_desc* d, int64_t P, int with_grad, int prec) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if (prec == COPE_PREC_BF16) return sdf_saved_floats_bf16(m, P, with_grad);
  return P * (m.pe_w + (int64_t)(m.n_lin - 1) * m.ldh * (2 + (with_grad ? 1 : 0)));
}

int64_t cope_sdf_ws_floats(const cope_mlp_desc* d, int64_t P, int prec) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if (prec == COPE_PREC_BF16) return sdf_ws_floats_bf16(m, P);
  int64_t ldw = (std::max(m.ldh, m.d_out) + 3) / 4 * 4;
  return P * ((int64_t)(m.n_lin + 6) * ldw + 4 * m.pe_w);
}

int64_t cope_sdf_query_ws_floats(const cope_mlp_desc* d, int64_t P, int prec) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if (prec == COPE_PREC_BF16) return sdf_query_ws_floats_bf16(m, P);
  return P * ((int64_t)m.pe_w + 3 * (int64_t)m.ldh) + 64;
}

int cope_sdf_query(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, float* sdf_out, float* ws,
                   int prec, cope_stream_t s_) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if ((prec & 0xFF) == COPE_PREC_BF16)
    return sdf_query_bf16(m, Wflat, x, P, sdf_out, ws, as_stream(s_), (prec & COPE_WS_HOLDS_PACK) != 0, flat_pack_ptr(m, Wflat, prec));
  COPE_REQUIRE(prec == COPE_PREC_FP32, "sdf_query: unknown precision %d", prec);
  if (P <= 0) return 0;
  cudaStream_t s = as_stream(s_);
  float* pe = ws;
  float* buf[2] = {pe + P * m.pe_w, pe + P * m.pe_w + P * m.ldh};
  float* skipbuf = buf[1] + P * m.ldh;   // dedicated input buffer of the skip layer
  pe_fwd_kernel<<<grid1d(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, 1, pe, m.pe_w,
                                                   m.skip > 0 ? skipbuf + (m.in[m.skip] - m.pe_w) : nullptr, m.ldh);
  COPE_CHECK_LAUNCH("pe_fwd");
  const float* in = pe;
  int ldin = m.pe_w;
  for (int l = 0; l < m.n_lin; ++l) {
    const bool last = l == m.n_lin - 1;
    float* out = last ? sdf_out : (l + 1 == m.skip ? skipbuf : buf[l & 1]);
    // the sdf-only query needs just row 0 of the last layer
    GemmArgs g = gemm_args((int)P, last ? 1 : m.out[l], m.in[l], in, ldin, Wflat + m.w_off[l], m.in[l], out,
                           last ? 1 : m.ldh);
    g.bias = Wflat + m.b_off[l];
    if (!last) { g.epi = EPI_BIAS_SOFTPLUS; g.alpha = (l + 1 == m.skip) ? kInvSqrt2 : 1.0f; g.act = m.act; g.act_slope = m.act_slope; }
    if (int rc = launch_gemm(false, true, g, s)) return rc;
    in = out; ldin = m.ldh;
  }
  return 0;
}

int cope_sdf_fwd(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, float* sdf, int sdf_ld,
                 float* feat, int feat_ld, float* grad, float* saved, float* ws, int prec, cope_stream_t s_) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if ((prec & 0xFF) == COPE_PREC_BF16)
    return sdf_fwd_bf16(m, Wflat, x, P, sdf, sdf_ld, feat, feat_ld, grad, saved, ws, as_stream(s_), nullptr, 0, false,
                        flat_pack_ptr(m, Wflat, prec));
  COPE_REQUIRE(prec == COPE_PREC_FP32, "sdf_fwd: unknown precision %d", prec);
  if (P <= 0) return 0;
  cudaStream_t s = as_stream(s_);
  SdfSaved sv = sdf_saved(m, P, saved);
  const int skw = m.skip > 0 ? m.in[m.skip] - m.pe_w : 0;   // width of the hidden part of the skip input
  pe_fwd_kernel<<<grid1d(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, 1, sv.pe, m.pe_w,
                                                   m.skip > 0 ? sv.h(m.skip) + skw : nullptr, m.ldh);
  COPE_CHECK_LAUNCH("pe_fwd");
  for (int l = 0; l < m.n_lin; ++l) {
    const bool last = l == m.n_lin - 1;
    GemmArgs g = gemm_args((int)P, m.out[l], m.in[l], sv.in(l), sv.ld_in(l), Wflat + m.w_off[l], m.in[l],
                           last ? sdf : sv.h(l + 1), last ? sdf_ld : m.ldh);
    g.bias = Wflat + m.b_off[l];
    if (!last) {
      g.epi = EPI_BIAS_SOFTPLUS; g.alpha = (l + 1 == m.skip) ? kInvSqrt2 : 1.0f; g.act = m.act; g.act_slope = m.act_slope;
      g.C2 = sv.z(l); g.ldc2 = m.ldh;
    } else {
      g.nsplit = 1; g.C2 = feat; g.ldc2 = feat_ld;
      if (!feat) g.N = 1;
    }
    if (int rc = launch_gemm(false, true, g, s)) return rc;
  }
  if (!grad) return 0;
  // ---- reverse sweep: grad = d y[:,0] / d x
  const int top = m.n_lin - 1;
  bcast_sigp_kernel<<<grid1d(P * m.out[top - 1]), 256, 0, s>>>(Wflat + m.w_off[top], sv.z(top - 1), m.ldh,
                                                              sv.dl(top - 1), m.ldh, P, m.out[top - 1], m.act, m.act_slope);
  COPE_CHECK_LAUNCH("bcast_sigp");
  float* ge0 = ws;                    // a_0            [P x pe_w]
  float* ge1 = ws + P * m.pe_w;       // a_skip pe part [P x pe_w]
  for (int l = top - 1; l >= 0; --l) {
    // a_l = delta_l * W_l  -> [P x in_l]
    GemmArgs g = gemm_args((int)P, m.in[l], m.out[l], sv.dl(l), m.ldh, Wflat + m.w_off[l], m.in[l],
                           l > 0 ? sv.dl(l - 1) : ge0, l > 0 ? m.ldh : m.pe_w);
    if (l > 0) {
      g.epi = EPI_MUL_SIGP; g.Z = sv.z(l - 1); g.ldz = m.ldh; g.act = m.act; g.act_slope = m.act_slope;
      if (l == m.skip) { g.alpha = kInvSqrt2; g.nsplit = skw; g.C2 = ge1; g.ldc2 = m.pe_w; }
    }
    if (int rc = launch_gemm(false, false, g, s)) return rc;
  }
  pe_vjp_kernel<<<grid1d(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, ge0, m.pe_w, m.skip > 0 ? ge1 : nullptr,
                                                   m.pe_w, grad, m.d_in, 0);
  COPE_CHECK_LAUNCH("pe_vjp");
  return 0;
}

int cope_sdf_bwd(const cope_mlp_desc* d, const float* Wflat, const float* x, int64_t P, const float* saved,
                 const float* d_sdf, int d_sdf_ld, const float* d_feat, int d_feat_ld, const float* dgrad,
                 float* dWflat, float* dx, int dx_accumulate, float* ws, int prec, cope_stream_t s_) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  COPE_REQUIRE(m.act == COPE_ACT_SOFTPLUS100 || (dgrad == nullptr && prec == COPE_PREC_FP32),
               "sdf_bwd: activation %d has no second-order / tensor-core path", m.act);
  if ((prec & 0xFF) == COPE_PREC_BF16)
    return sdf_bwd_bf16(m, Wflat, x, P, saved, d_sdf, d_sdf_ld, d_feat, d_feat_ld, dgrad, dWflat, dx, dx_accumulate, ws,
                        as_stream(s_), false, flat_pack_ptr(m, Wflat, prec));
  COPE_REQUIRE(prec == COPE_PREC_FP32, "sdf_bwd: unknown precision %d", prec);
  const bool have_dy = d_sdf || d_feat;
  if (P <= 0 || (!have_dy && !dgrad)) {
    if (dx && P > 0 && !dx_accumulate) cudaMemsetAsync(dx, 0, sizeof(float) * P * m.d_in, as_stream(s_));
    return 0;
  }
  cudaStream_t s = as_stream(s_);
  SdfSaved sv = sdf_saved(m, P, const_cast<float*>(saved));
  const int top = m.n_lin - 1;
  const int skw = m.skip > 0 ? m.in[m.skip] - m.pe_w : 0;
  const int64_t ldw = (std::max(m.ldh, m.d_out) + 3) / 4 * 4;
  // workspace carve-up
  float* T[2] = {ws, ws + P * ldw};
  float* ZB2 = T[1] + P * ldw;                      // (n_lin-1) x [P x ldh]
  float* ZB[2] = {ZB2 + (int64_t)(m.n_lin - 1) * P * m.ldh, nullptr};
  ZB[1] = ZB[0] + P * ldw;
  float* t0 = ZB[1] + P * ldw;                      // [P x pe_w]
  float* eb0 = t0 + P * m.pe_w;                     // [P x pe_w]
  float* eb1 = eb0 + P * m.pe_w;                    // [P x pe_w]
  float* dyb = eb1 + P * m.pe_w;                    // [P x ldw] packed adjoint of the last layer
  const float* dy = nullptr;
  if (have_dy) {
    pack_dy_kernel<<<grid1d(P * m.d_out), 256, 0, s>>>(d_sdf, d_sdf_ld, d_feat, d_feat_ld, P, m.d_out, dyb, (int)ldw);
    COPE_CHECK_LAUNCH("pack_dy");
    dy = dyb;
  }
  auto zb2 = [&](int l) { return ZB2 + (int64_t)l * P * m.ldh; };

  if (dgrad) {
    // ---- tangent pass (the double backward) + second-order weight gradients
    // T_l lives in T[l&1] except T_0 (t0) ; the skip input's pe part is t0/sqrt2
    float* tskip = m.skip > 0 ? T[m.skip & 1] : nullptr;
    pe_jvp_kernel<<<grid1d(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, dgrad, t0, m.pe_w, nullptr, 0);
    COPE_CHECK_LAUNCH("pe_jvp");
    for (int l = 0; l < top; ++l) {
      const float* tin = l == 0 ? t0 : T[l & 1];
      const int ldt = l == 0 ? m.pe_w : (int)ldw;
      if (l == m.skip) {
        strided_copy_kernel<<<grid1d(P * m.pe_w), 256, 0, s>>>(t0, m.pe_w, tskip + skw, (int)ldw, P, m.pe_w, kInvSqrt2);
        COPE_CHECK_LAUNCH("skip_copy");
      }
      if (int rc = wgrad(sv.dl(l), m.ldh, tin, ldt, P, m.out[l], m.in[l], dWflat + m.w_off[l], s)) return rc;
      GemmArgs g = gemm_args((int)P, m.out[l], m.in[l], tin, ldt, Wflat + m.w_off[l], m.in[l], T[(l + 1) & 1], (int)ldw);
      g.epi = EPI_TANGENT; g.alpha = (l + 1 == m.skip) ? kInvSqrt2 : 1.0f;
      g.Z = sv.z(l); g.ldz = m.ldh; g.D = sv.dl(l); g.ldd = m.ldh; g.C2 = zb2(l); g.ldc2 = m.ldh;
      if (int rc = launch_gemm(false, true, g, s)) return rc;
    }
    // last layer: delta_top = e_0  =>  dW_top[0, :] += sum_p t_top[p, :]
    if (int rc = colsum(T[top & 1], (int)ldw, P, m.in[top], dWflat + m.w_off[top], s)) return rc;
  }

  // ---- adjoint sweep(s).  pass 0: weights (value + second-order adjoints); pass 1 (only if dx is wanted
  // and a second-order term exists): value path only, data GEMMs only.
  const int n_pass = (dx && dgrad) ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    const bool do_w = pass == 0;
    const bool with2 = dgrad && pass == 0;
    const bool want_e = dx && (pass == n_pass - 1);
    if (!dy && !with2) break;
    const float* zb = dy;   // adjoint of z_top, [P x d_out]
    int ldzb = (int)ldw;
    int l_start = top;
    if (!dy) { zb = zb2(top - 1); ldzb = m.ldh; l_start = top - 1; }
    for (int l = l_start; l >= 0; --l) {
      if (do_w) {
        if (int rc = wgrad(zb, ldzb, sv.in(l), sv.ld_in(l), P, m.out[l], m.in[l], dWflat + m.w_off[l], s)) return rc;
        if (int rc = colsum(zb, ldzb, P, m.out[l], dWflat + m.b_off[l], s)) return rc;
      }
      if (l == 0 && !want_e) break;
      float* nxt = ZB[l & 1];
      GemmArgs g = gemm_args((int)P, m.in[l], m.out[l], zb, ldzb, Wflat + m.w_off[l], m.in[l], l > 0 ? nxt : eb0,
                             l > 0 ? (int)ldw : m.pe_w);
      if (l > 0) {
        g.epi = EPI_BWD; g.Z = sv.z(l - 1); g.ldz = m.ldh; g.act = m.act; g.act_slope = m.act_slope;
        if (with2) { g.D = zb2(l - 1); g.ldd = m.ldh; }
        if (l == m.skip) { g.alpha = kInvSqrt2; g.nsplit = skw; g.C2 = want_e ? eb1 : nullptr; g.ldc2 = m.pe_w; }
      }
      if (int rc = launch_gemm(false, false, g, s)) return rc;
      zb = nxt; ldzb = (int)ldw;
    }
    if (want_e) {
      pe_vjp_kernel<<<grid1d(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, eb0, m.pe_w,
                                                       (m.skip > 0 && l_start >= m.skip) ? eb1 : nullptr, m.pe_w, dx,
                                                       m.d_in, dx_accumulate);
      COPE_CHECK_LAUNCH("pe_vjp");
    }
  }
  if (dx && !dy && !dx_accumulate) cudaMemsetAsync(dx, 0, sizeof(float) * P * m.d_in, s);
  return 0;
}


// ================================================================================================ colour
// input row = [x(4) | PE_Lv(dirs) | normals(4) | feat]; saved = [cin | H_1..H_{n-1} (post-ReLU) | rgb]
}  // extern "C"

namespace cope {
__global__ void color_pack_kernel(const float* __restrict__ x, const float* __restrict__ dirs, int dirs_group, int Lv,
                                  const float* __restrict__ nrm, const float* __restrict__ feat, int feat_ld, int d_feat,
                                  int64_t P, float* __restrict__ cin, int ld) {
  const int pe_w = 3 * (1 + 2 * Lv);
  const int head = 4 + pe_w + 4;
  const int W = head + d_feat;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * W) return;
  int64_t p = i / W;
  int c = (int)(i - p * W);
  float v;
  if (c < 4) v = x[p * 4 + c];
  else if (c < 4 + pe_w) {
    int e = c - 4;
    const float* dv = dirs + (p / dirs_group) * 3;
    if (e < 3) v = dv[e];
    else {
      int blk = (e - 3) / 3, dd = (e - 3) % 3;     // blk = 2k (sin) or 2k+1 (cos)
      float a = dv[dd] * (float)(1 << (blk >> 1));
      v = (blk & 1) ? cosf(a) : sinf(a);
    }
  } else if (c < head) v = nrm[p * 4 + (c - 4 - pe_w)];
  else v = feat[p * feat_ld + (c - head)];
  cin[p * ld + c] = v;
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, float* __restrict__ dz,
                                   int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float o = out[i]; dz[i] = dout[i] * o * (1.0f - o); }
}

// scatter d_cin [P x W] back to the four inputs
__global__ void color_unpack_kernel(const float* __restrict__ dcin, int ld, const float* __restrict__ dirs, int dirs_group,
                                    int Lv, int d_feat, int64_t P, float* __restrict__ dx, float* __restrict__ ddirs,
                                    float* __restrict__ dnrm, float* __restrict__ dfeat, int dfeat_ld) {
  const int pe_w = 3 * (1 + 2 * Lv);
  const int head = 4 + pe_w + 4;
  const int W = 4 + 3 + 4 + d_feat;   // work items per point: dx(4), ddirs(3), dnrm(4), dfeat
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * W) return;
  int64_t p = i / W;
  int c = (int)(i - p * W);
  const float* r = dcin + p * ld;
  if (c < 4) { if (dx) dx[p * 4 + c] += r[c]; }
  else if (c < 7) {
    if (!ddirs) return;
    int dd = c - 4;
    float v = dirs[(p / dirs_group) * 3 + dd];
    const float* e = r + 4;
    float acc = e[dd], f = 1.0f;
    for (int k = 0; k < Lv; ++k, f *= 2.0f) {
      float sn, cs;
      sincosf(v * f, &sn, &cs);
      acc += f * (cs * e[3 * (1 + 2 * k) + dd] - sn * e[3 * (2 + 2 * k) + dd]);
    }
    ddirs[p * 3 + dd] = acc;
  } else if (c < 11) { if (dnrm) dnrm[p * 4 + (c - 7)] += r[4 + pe_w + (c - 7)]; }
  else if (dfeat) dfeat[p * dfeat_ld + (c - 11)] = r[head + (c - 11)];
}

struct ColorSaved {
  float* cin; float* H; float* rgb; int64_t P; int ldh, w0;
  float* h(int l) const { return H + (int64_t)(l - 1) * P * ldh; }   // input of layer l >= 1
  const float* in(int l) const { return l == 0 ? cin : h(l); }
  int ld_in(int l) const { return l == 0 ? w0 : ldh; }
};
static ColorSaved color_saved(const MlpShape& m, int64_t P, float* base) {
  ColorSaved v; v.P = P; v.ldh = m.ldh; v.w0 = m.in[0];
  v.cin = base; v.H = base + P * m.in[0]; v.rgb = v.H + (int64_t)(m.n_lin - 1) * P * m.ldh;
  return v;
}
}  // namespace cope

extern "C" {

int64_t cope_color_saved_floats(const cope_mlp_desc* d, int64_t P, int prec) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if (prec == COPE_PREC_BF16) return color_saved_floats_bf16(m, P);
  return P * (m.in[0] + (int64_t)(m.n_lin - 1) * m.ldh + m.d_out);
}
int64_t cope_color_ws_floats(const cope_mlp_desc* d, int64_t P, int prec) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if (prec == COPE_PREC_BF16) return color_ws_floats_bf16(m, P);
  return P * (2 * (int64_t)m.ldh + m.in[0] + 4 + m.d_out);
}

int cope_color_fwd(const cope_mlp_desc* d, const float* Wflat, const float* x, const float* dirs, int dirs_group, int Lv,
                   const float* normals, const float* feat, int feat_ld, int64_t P, float* rgb, float* saved, float* ws,
                   int prec, cope_stream_t s_) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if ((prec & 0xFF) == COPE_PREC_BF16)
    return color_fwd_bf16(m, Wflat, x, dirs, dirs_group, Lv, normals, feat, feat_ld, P, rgb, saved, ws, as_stream(s_), false, false,
                          flat_pack_ptr(m, Wflat, prec));
  COPE_REQUIRE(prec == COPE_PREC_FP32, "color_fwd: unknown precision %d", prec);
  const int d_feat = m.in[0] - (4 + 3 * (1 + 2 * Lv) + 4);
  COPE_REQUIRE(d_feat > 0 && m.skip < 0, "color_fwd: layer-0 width %d does not match idr input", m.in[0]);
  if (P <= 0) return 0;
  (void)ws;
  cudaStream_t s = as_stream(s_);
  ColorSaved sv = color_saved(m, P, saved);
  color_pack_kernel<<<grid1d(P * m.in[0]), 256, 0, s>>>(x, dirs, dirs_group, Lv, normals, feat, feat_ld, d_feat, P, sv.cin,
                                                       m.in[0]);
  COPE_CHECK_LAUNCH("color_pack");
  for (int l = 0; l < m.n_lin; ++l) {
    const bool last = l == m.n_lin - 1;
    GemmArgs g = gemm_args((int)P, m.out[l], m.in[l], sv.in(l), sv.ld_in(l), Wflat + m.w_off[l], m.in[l],
                           last ? sv.rgb : sv.h(l + 1), last ? m.d_out : m.ldh);
    g.bias = Wflat + m.b_off[l];
    g.epi = last ? EPI_BIAS_SIGMOID : EPI_BIAS_RELU;
    if (int rc = launch_gemm(false, true, g, s)) return rc;
  }
  cudaMemcpyAsync(rgb, sv.rgb, sizeof(float) * P * m.d_out, cudaMemcpyDeviceToDevice, s);
  return 0;
}

int cope_color_bwd(const cope_mlp_desc* d, const float* Wflat, const float* dirs, int dirs_group, int Lv, int64_t P,
                   const float* saved, const float* d_rgb, float* dWflat, float* dx, float* ddirs, float* dnormals,
                   float* dfeat, int dfeat_ld, float* ws, int prec, cope_stream_t s_) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  if ((prec & 0xFF) == COPE_PREC_BF16)
    return color_bwd_bf16(m, Wflat, dirs, dirs_group, Lv, P, saved, d_rgb, dWflat, dx, ddirs, dnormals, dfeat, dfeat_ld, ws,
                          as_stream(s_), nullptr, 0, flat_pack_ptr(m, Wflat, prec));
  COPE_REQUIRE(prec == COPE_PREC_FP32, "color_bwd: unknown precision %d", prec);
  const int d_feat = m.in[0] - (4 + 3 * (1 + 2 * Lv) + 4);
  if (P <= 0) return 0;
  cudaStream_t s = as_stream(s_);
  ColorSaved sv = color_saved(m, P, const_cast<float*>(saved));
  float* B[2] = {ws, ws + P * m.ldh};
  float* dcin = B[1] + P * m.ldh;
  float* dzl = dcin + P * m.in[0];
  sigmoid_bwd_kernel<<<grid1d(P * m.d_out), 256, 0, s>>>(d_rgb, sv.rgb, dzl, P * m.d_out);
  COPE_CHECK_LAUNCH("sigmoid_bwd");
  const float* dz = dzl;
  int lddz = m.d_out;
  for (int l = m.n_lin - 1; l >= 0; --l) {
    if (int rc = wgrad(dz, lddz, sv.in(l), sv.ld_in(l), P, m.out[l], m.in[l], dWflat + m.w_off[l], s)) return rc;
    if (int rc = colsum(dz, lddz, P, m.out[l], dWflat + m.b_off[l], s)) return rc;
    float* nxt = l > 0 ? B[l & 1] : dcin;
    GemmArgs g = gemm_args((int)P, m.in[l], m.out[l], dz, lddz, Wflat + m.w_off[l], m.in[l], nxt, l > 0 ? m.ldh : m.in[0]);
    if (l > 0) { g.epi = EPI_RELU_MASK; g.Z = sv.h(l); g.ldz = m.ldh; }
    if (int rc = launch_gemm(false, false, g, s)) return rc;
    dz = nxt; lddz = m.ldh;
  }
  color_unpack_kernel<<<grid1d(P * (11 + d_feat)), 256, 0, s>>>(dcin, m.in[0], dirs, dirs_group, Lv, d_feat, P, dx, ddirs,
                                                                dnormals, dfeat, dfeat_ld);
  COPE_CHECK_LAUNCH("color_unpack");
  return 0;
}

/* ---- weights packed once per step (bf16 path) */
int64_t cope_mlp_pack_offset(const cope_mlp_desc* d) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  return (m.n_flat + 63) / 64 * 64;
}
int64_t cope_mlp_pack_floats(const cope_mlp_desc* d, int is_color, int Lv) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  const int64_t e = mlp_pack_elems_bf16(m, is_color, Lv);
  return e < 0 ? -1 : (e + 1) / 2 + 64;
}
int cope_mlp_pack(const cope_mlp_desc* d, int is_color, int Lv, float* flat_with_tail, cope_stream_t s) {
  MlpShape m;
  if (make_shape(d, &m)) return -1;
  return mlp_pack_bf16(m, is_color, Lv, flat_with_tail, reinterpret_cast<__nv_bfloat16*>(flat_with_tail + (m.n_flat + 63) / 64 * 64),
                       as_stream(s));
}

}  // extern "C"
