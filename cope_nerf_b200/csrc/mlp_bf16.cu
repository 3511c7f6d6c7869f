// SDF / colour MLPs on the bf16 tcgen05 path (COPE_PREC_BF16): same algebra as mlp_f32.cu, every dense layer is a
// tc_gemm launch (weights packed to bf16 once per call), every weight gradient a tc_wgrad launch.
//
// Precision layout
//   * activations / adjoints between layers: bf16 [P x LD]; accumulation fp32 in TMEM
//   * softplus'(z) is recovered from the stored activation h = softplus(z) as 1 - exp(-100 h): accurate exactly where
//     the derivative is sensitive (small h), so pre-activations are never stored
//   * the raw coordinates enter layer 0 / the colour net as hi + lo bf16 pairs (duplicated weight columns), so the
//     position is not quantised to 8 bits; everything that touches the PE Jacobian (ge0/ge1, dx, d_dirs) stays fp32
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "mlp_shape.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"
#include "sdf_fused.cuh"

namespace cope {

// sdf_chain.cu: fully fused forward chain (reference architecture only)
bool sdf_chain_supported(const MlpShape& m);
int launch_sdf_chain_query(const MlpShape& m, const float* Wflat, const bf16* wp, const uint32_t* wf_off, uint32_t wtop_off,
                           const float* x, int64_t P, float* sdf_out, cudaStream_t s);

static inline int r16(int v) { return (v + 15) / 16 * 16; }
static inline int r64(int v) { return (v + 63) / 64 * 64; }
static inline int r128(int v) { return (v + 127) / 128 * 128; }
static inline dim3 g1(int64_t n, int bs = 256) { return dim3((unsigned)ceil_div(n, bs)); }

// ------------------------------------------------------------------------------------------- small kernels
// pe[p, :] = [x_hi | sin/cos | x_lo | 0];  optional skip copy (x_hi | sin/cos)/sqrt2 into dst2 (bf16)
__global__ void pe_fwd_bf16_kernel(const float* __restrict__ x, int64_t P, int d, int L, bf16* __restrict__ out, int ld,
                                   int width, bf16* __restrict__ out2, int ld2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d) return;
  int64_t p = i / d;
  int dd = (int)(i - p * d);
  const int pe_w = d * (1 + 2 * L);
  float v = x[p * d + dd];
  bf16 hi = __float2bfloat16(v);
  bf16* o = out + p * ld;
  o[dd] = hi;
  o[pe_w + dd] = __float2bfloat16(v - __bfloat162float(hi));
  if (dd == 0) for (int c = pe_w + d; c < width; ++c) o[c] = __float2bfloat16(0.0f);
  bf16* o2 = out2 ? out2 + p * ld2 : nullptr;
  if (o2) o2[dd] = __float2bfloat16(v * kInvSqrt2);
  float f = 1.0f;
  for (int k = 0; k < L; ++k, f *= 2.0f) {
    float s, c;
    sincosf(v * f, &s, &c);
    o[d * (1 + 2 * k) + dd] = __float2bfloat16(s);
    o[d * (2 + 2 * k) + dd] = __float2bfloat16(c);
    if (o2) {
      o2[d * (1 + 2 * k) + dd] = __float2bfloat16(s * kInvSqrt2);
      o2[d * (2 + 2 * k) + dd] = __float2bfloat16(c * kInvSqrt2);
    }
  }
}

// t0[p, :] = [J_PE G | 0]  (bf16), optional skip copy / sqrt2
__global__ void pe_jvp_bf16_kernel(const float* __restrict__ x, int64_t P, int d, int L, const float* __restrict__ G,
                                   bf16* __restrict__ t0, int ld, int width, bf16* __restrict__ t2, int ld2) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * d) return;
  int64_t p = i / d;
  int dd = (int)(i - p * d);
  const int pe_w = d * (1 + 2 * L);
  float v = x[p * d + dd], gv = G[p * d + dd];
  bf16* o = t0 + p * ld;
  bf16* o2 = t2 ? t2 + p * ld2 : nullptr;
  o[dd] = __float2bfloat16(gv);
  if (dd == 0) for (int c = pe_w; c < width; ++c) o[c] = __float2bfloat16(0.0f);
  if (o2) o2[dd] = __float2bfloat16(gv * kInvSqrt2);
  float f = 1.0f;
  for (int k = 0; k < L; ++k, f *= 2.0f) {
    float s, c;
    sincosf(v * f, &s, &c);
    float ts = f * c * gv, tc = -f * s * gv;
    o[d * (1 + 2 * k) + dd] = __float2bfloat16(ts);
    o[d * (2 + 2 * k) + dd] = __float2bfloat16(tc);
    if (o2) {
      o2[d * (1 + 2 * k) + dd] = __float2bfloat16(ts * kInvSqrt2);
      o2[d * (2 + 2 * k) + dd] = __float2bfloat16(tc * kInvSqrt2);
    }
  }
}

// D[p, n] = w[n] * (1 - exp(-100 * H[p, n] * hscale))    (top of the reverse sweep), pad columns zeroed; 8 cols / thread
__global__ void bcast_sp_bf16_kernel(const float* __restrict__ w, const bf16* __restrict__ H, int ldh, float hscale,
                                     bf16* __restrict__ D, int ldd, int64_t P, int n, int npad) {
  const int groups = npad >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * groups) return;
  const int64_t p = i / groups;
  const int c0 = (int)(i - p * groups) * 8;
  float v[8];
  const uint4 hq = *reinterpret_cast<const uint4*>(H + p * ldh + c0);
  const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float h = (k & 1) ? tc::bf16_hi(hw[k >> 1]) : tc::bf16_lo(hw[k >> 1]);
    v[k] = (c0 + k < n) ? w[c0 + k] * (1.0f - __expf(-kSoftplusBeta * h * hscale)) : 0.0f;
  }
  uint4 q;
  q.x = tc::pack_bf16(v[0], v[1]); q.y = tc::pack_bf16(v[2], v[3]); q.z = tc::pack_bf16(v[4], v[5]); q.w = tc::pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(D + p * ldd + c0) = q;
}

// out[c] += sum_p w[p*ldw] * X[p, c]   (w == null -> 1), n <= 256 columns (n % 8 == 0).  One warp reads whole 512-byte rows
// (16 bytes per lane), 8 warps per block take interleaved rows, 4 rows in flight per warp; block partials meet in shared
// memory and leave as one atomic per column.
__global__ void __launch_bounds__(256) wcolsum_bf16_kernel(const bf16* __restrict__ X, int ld, const float* __restrict__ w, int ldw,
                                                           int64_t P, int n, int rows_per_block, float* __restrict__ out) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  const int c0 = lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < n) {
    int64_t p = p0 + warp;
    for (; p + 24 < p1; p += 32) {
      uint4 q[4];
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        q[u] = __ldcs(reinterpret_cast<const uint4*>(X + (p + 8 * u) * ld + c0));
        wv[u] = w ? w[(p + 8 * u) * ldw] : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t ww[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] = fmaf(wv[u], tc::bf16_lo(ww[k]), acc[2 * k]);
          acc[2 * k + 1] = fmaf(wv[u], tc::bf16_hi(ww[k]), acc[2 * k + 1]);
        }
      }
    }
    for (; p < p1; p += 8) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(X + p * ld + c0));
      const float wv = w ? w[p * ldw] : 1.0f;
      const uint32_t ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] = fmaf(wv, tc::bf16_lo(ww[k]), acc[2 * k]);
        acc[2 * k + 1] = fmaf(wv, tc::bf16_hi(ww[k]), acc[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][c0 + k] = acc[k];
  __syncthreads();
  const int c = threadIdx.x;
  if (c < n) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][c];
    atomicAdd(out + c, s);
  }
}
__global__ void colsum_f32_strided_kernel(const float* __restrict__ X, int ld, int64_t P, int n, int rows_per_block,
                                          float* __restrict__ out) {
  int64_t p0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t p1 = p0 + rows_per_block < P ? p0 + rows_per_block : P;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    float acc = 0.0f;
    for (int64_t p = p0; p < p1; ++p) acc += X[p * ld + c];
    atomicAdd(out + c, acc);
  }
}

// dst[p, c] = bf16(src[p*lds + c] * scale) for c < n, 0 for n <= c < npad   (8 columns per thread, npad % 8 == 0)
__global__ void cvt_f32_bf16_kernel(const float* __restrict__ src, int lds, bf16* __restrict__ dst, int ldd, int64_t P, int n,
                                    int npad, float scale) {
  const int groups = npad >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * groups) return;
  const int64_t p = i / groups;
  const int c0 = (int)(i - p * groups) * 8;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = (src && c0 + k < n) ? src[p * lds + c0 + k] * scale : 0.0f;
  uint4 q;
  q.x = tc::pack_bf16(v[0], v[1]); q.y = tc::pack_bf16(v[2], v[3]); q.z = tc::pack_bf16(v[4], v[5]); q.w = tc::pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst + p * ldd + c0) = q;
}
// out[0] += sum_p x[p * ld]
__global__ void sum_strided_kernel(const float* __restrict__ x, int ld, int64_t P, float* __restrict__ out) {
  float acc = 0.0f;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) acc += x[p * ld];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
__global__ void copy_bf16_scaled_kernel(const bf16* __restrict__ src, int lds, bf16* __restrict__ dst, int ldd, int64_t P,
                                        int n, float scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * n) return;
  int64_t p = i / n;
  int c = (int)(i - p * n);
  dst[p * ldd + c] = __float2bfloat16(__bfloat162float(src[p * lds + c]) * scale);
}
__global__ void zero_cols_bf16_kernel(bf16* __restrict__ dst, int ld, int64_t P, int c0, int c1) {
  const int w = c1 - c0;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * w) return;
  int64_t p = i / w;
  dst[p * ld + c0 + (int)(i - p * w)] = __float2bfloat16(0.0f);
}

static int wcolsum(const bf16* X, int ld, const float* w, int ldw, int64_t P, int n, float* out, cudaStream_t s) {
  if (P <= 0 || n <= 0) return 0;
  COPE_REQUIRE(n <= 256 && n % 8 == 0 && ld % 8 == 0, "wcolsum: n=%d (<= 256, multiple of 8) ld=%d (multiple of 8)", n, ld);
  const int rows = 256;
  wcolsum_bf16_kernel<<<(unsigned)ceil_div(P, rows), 256, 0, s>>>(X, ld, w, ldw, P, n, rows, out);
  COPE_CHECK_LAUNCH("wcolsum_bf16");
  return 0;
}

// ------------------------------------------------------------------------------------------- SDF layout
struct SdfB {
  int n_lin, top, skip, skw, pe_w, pe_k, LD, d_in, L, d_out, featN;
  int Kp[COPE_MAX_LIN], Np[COPE_MAX_LIN];
  size_t wf_off[COPE_MAX_LIN], wt_off[COPE_MAX_LIN], w_total;   // element offsets of packed W_l / W_l^T
  size_t wtop_sdf_off;                                          // packed 16-row block holding row 0 of the last layer
};
static int make_sdfb(const MlpShape& m, SdfB* b) {
  COPE_REQUIRE(m.act == COPE_ACT_SOFTPLUS100, "bf16 path: only the softplus(beta=100) SDF network is supported (activation %d)", m.act);
  b->n_lin = m.n_lin; b->top = m.n_lin - 1; b->skip = m.skip; b->pe_w = m.pe_w; b->d_in = m.d_in; b->L = m.L;
  b->skw = m.skip > 0 ? m.in[m.skip] - m.pe_w : 0;
  b->pe_k = r64(m.pe_w + m.d_in);
  b->d_out = m.d_out; b->featN = r16(m.d_out - 1);
  COPE_REQUIRE(b->pe_k == 64, "bf16 path: PE width %d + %d does not fit one 64-wide K chunk", m.pe_w, m.d_in);
  COPE_REQUIRE(m.d_out - 1 >= 1 && b->featN <= 256, "bf16 path: d_out-1=%d must be in [1,256]", m.d_out - 1);
  int LD = 128;   // >= 128 so every adjoint buffer can be the 128-row-padded X operand of tc_wgrad
  size_t off = 0;
  for (int l = 0; l < m.n_lin; ++l) {
    b->Kp[l] = l == 0 ? b->pe_k : r64(m.in[l]);
    b->Np[l] = r16(m.out[l]);
    if (l > 0) {
      COPE_REQUIRE(m.in[l] % 64 == 0 && m.in[l] <= 256, "bf16 path: hidden width %d must be a multiple of 64, <= 256", m.in[l]);
      LD = std::max(LD, b->Kp[l]);
    }
    const int npf = l == b->top ? b->featN : b->Np[l];
    b->wf_off[l] = off; off += (size_t)npf * b->Kp[l];
    // transposed: B[n = input col][k = output row]
    const int nt = l == 0 ? 64 : r16(m.in[l]);
    const int kt = l == b->top ? r64(m.d_out - 1) : r64(m.out[l]);
    b->wt_off[l] = off; off += (size_t)nt * kt;
  }
  b->wtop_sdf_off = off; off += (size_t)16 * b->Kp[b->top];
  b->w_total = (off + 63) / 64 * 64;
  b->LD = LD;
  return 0;
}

// pack every layer (forward always; transposed when asked)
static int pack_sdf(const MlpShape& m, const SdfB& b, const float* Wflat, bf16* wp, bool fwd, bool transposed, cudaStream_t s) {
  PackBatch pb{};
  for (int l = 0; l < m.n_lin; ++l) {
    const float* W = Wflat + m.w_off[l];
    const bool top = l == b.top;
    if (fwd) {
      PackSpec sp = pack_spec();
      if (top) seg_n(sp, 0, 1, m.d_out - 1); else seg_n(sp, 0, 0, m.out[l]);
      if (l == 0) { seg_k(sp, 0, 0, m.pe_w); seg_k(sp, m.pe_w, 0, m.d_in); } else seg_k(sp, 0, 0, m.in[l]);
      pack_add(pb, W, m.in[l], sp, top ? b.featN : b.Np[l], b.Kp[l], 0, wp + b.wf_off[l]);
      if (top) {
        PackSpec s0 = pack_spec();
        seg_n(s0, 0, 0, 1); seg_k(s0, 0, 0, m.in[l]);
        pack_add(pb, W, m.in[l], s0, 16, b.Kp[l], 0, wp + b.wtop_sdf_off);
      }
    }
    if (transposed) {
      PackSpec sp = pack_spec();
      seg_n(sp, 0, 0, l == 0 ? m.pe_w : m.in[l]);
      if (top) seg_k(sp, 0, 1, m.d_out - 1); else seg_k(sp, 0, 0, m.out[l]);
      const int nt = l == 0 ? 64 : r16(m.in[l]);
      const int kt = top ? r64(m.d_out - 1) : r64(m.out[l]);
      pack_add(pb, W, m.in[l], sp, nt, kt, 1, wp + b.wt_off[l]);
    }
  }
  return launch_tc_pack_batch(pb, s);
}

struct SdfSavedB {
  bf16* pe; bf16* H; bf16* D; int64_t P; int LD;
  bf16* h(int l) const { return H + (int64_t)(l - 1) * P * LD; }   // input of layer l >= 1
  bf16* dl(int l) const { return D + (int64_t)l * P * LD; }        // delta_l, l < top
  const bf16* in(int l) const { return l == 0 ? pe : h(l); }
  int ld_in(int l) const { return l == 0 ? 64 : LD; }
};
static SdfSavedB sdf_saved_b(const SdfB& b, int64_t P, float* base) {
  SdfSavedB v; v.P = P; v.LD = b.LD;
  v.pe = reinterpret_cast<bf16*>(base);
  v.H = v.pe + P * 64;
  v.D = v.H + (int64_t)b.top * P * b.LD;
  return v;
}
static inline float hscale_of(const SdfB& b, int l) { return l == b.skip ? 1.41421356237309505f : 1.0f; }   // H_l = alpha_l * softplus
static inline float alpha_of(const SdfB& b, int l) { return l == b.skip ? kInvSqrt2 : 1.0f; }

int64_t sdf_saved_floats_bf16(const MlpShape& m, int64_t P, int with_grad) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  int64_t elems = P * (64 + (int64_t)b.top * b.LD * (with_grad ? 2 : 1));
  return (elems + 1) / 2 + 64;
}
int64_t sdf_ws_floats_bf16(const MlpShape& m, int64_t P) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  // packed weights + (T: top, ZB2: top, ZB: top, dyb: 1, spare: 1) bf16 [P x LD] + t0 bf16 [P x 64] + 4 fp32 [P x 64]
  int64_t bf = (int64_t)b.w_total + P * ((int64_t)(3 * b.top + 2) * b.LD + 64);
  return (bf + 1) / 2 + P * 4 * 64 + tc_wgrad_part_floats() + 1024;
}

// workspace of the sdf-only query: packed weights (the fused chain needs nothing else); the layer-by-layer fallback adds the
// PE and three activation buffers
int64_t sdf_query_ws_floats_bf16(const MlpShape& m, int64_t P) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  int64_t bf = (int64_t)b.w_total;
  if (!sdf_chain_supported(m) || getenv("COPE_NO_CHAIN")) bf += P * (64 + 3 * (int64_t)b.LD);
  return (bf + 1) / 2 + 1024;
}

// ------------------------------------------------------------------------------------------- SDF forward
static int sdf_layers_fwd(const MlpShape& m, const SdfB& b, const float* Wflat, const bf16* wp, const float* x, int64_t P,
                          const bf16* pe, bf16* const* hbuf /* hbuf[l] = input buffer of layer l (l>=1) */, float* sdf,
                          int sdf_ld, float* feat, int feat_ld, cudaStream_t s, bf16* feat_b16 = nullptr, int feat_b16_ld = 0) {
  for (int l = 0; l < b.top; ++l) {
    TcArgs t = tc_args((int)P, b.Np[l], b.Kp[l], l == 0 ? pe : hbuf[l], l == 0 ? 64 : b.LD, wp + b.wf_off[l], hbuf[l + 1], b.LD, 0);
    t.bias = Wflat + m.b_off[l]; t.epi = TC_BIAS_SOFTPLUS; t.alpha = alpha_of(b, l + 1); t.n_valid = m.out[l];
    if (int rc = launch_tc_gemm(t, s)) return rc;
  }
  const int l = b.top;
  if (feat || feat_b16) {
    TcArgs t = feat_b16 ? tc_args((int)P, b.featN, b.Kp[l], hbuf[l], b.LD, wp + b.wf_off[l], feat_b16, feat_b16_ld, 0)
                        : tc_args((int)P, b.featN, b.Kp[l], hbuf[l], b.LD, wp + b.wf_off[l], feat, feat_ld, 1);
    t.bias = Wflat + m.b_off[l] + 1; t.n_valid = m.d_out - 1;
    if (int rc = launch_tc_gemm(t, s)) return rc;
  }
  TcArgs t = tc_args((int)P, 16, b.Kp[l], hbuf[l], b.LD, wp + b.wtop_sdf_off, sdf, sdf_ld, 1);
  t.bias = Wflat + m.b_off[l]; t.n_valid = 1;
  return launch_tc_gemm(t, s);
}


// ------------------------------------------------------------------------------------------- fused chains (sdf_fused.cu)
// COPE_NO_FUSED=1 (or "all") disables every fused chain; a list such as "fwd,adj" disables single passes (A/B tests)
static bool fused_enabled(const MlpShape& m, const SdfB& b, const char* pass) {
  const char* e = getenv("COPE_NO_FUSED");
  if (e && (strstr(e, pass) || strstr(e, "all") || !strcmp(e, "1"))) return false;
  return b.LD == 256 && sdf_fused_supported(m);
}
// true when the inference forward (sdf_fwd_bf16 with infer) runs on the fused chain and therefore never writes the delta
// stack: only then may the caller size the saved block with with_grad = 0 (the layer-by-layer fallback still stores delta_l)
bool sdf_infer_compact_bf16(const MlpShape& m) {
  SdfB b;
  if (make_sdfb(m, &b)) return false;
  return fused_enabled(m, b, "fwd");
}
static void fz_common(const MlpShape& m, const SdfB& b, const float* Wflat, const bf16* wp, const float* x, int64_t P, FzArgs* a) {
  a->P = P; a->x = x; a->Wflat = Wflat; a->wp = wp;
  a->n_lin = m.n_lin; a->skip = m.skip; a->skw = b.skw; a->pe_w = m.pe_w; a->d_in = m.d_in; a->L = m.L;
  for (int l = 0; l < m.n_lin; ++l) a->b_off[l] = m.b_off[l];
  a->w_top_off = m.w_off[b.top];
}
static void fz_job(FzArgs* a, size_t w_off, int Np, int Kp, int acc, int wait_a, int commit) {
  FzJob& j = a->jobs[a->n_jobs++];
  j.w_off = (uint32_t)w_off; j.Np = (uint16_t)Np; j.Kp = (uint16_t)Kp; j.acc = (uint8_t)acc; j.wait_a = (uint8_t)wait_a;
  j.commit = (uint8_t)commit; j.pad = 0;
}

// value pass + reverse sweep in ONE kernel: writes sv.pe, sv.h(1..top), sv.dl(0..top-1), sdf, the bf16 feature block and
// ge0 / ge1 (gradient w.r.t. the PE)
static int sdf_fwd_fused(const MlpShape& m, const SdfB& b, const float* Wflat, const bf16* wp, const float* x, int64_t P,
                         const SdfSavedB& sv, float* sdf, int sdf_ld, bf16* feat_b16, int feat_b16_ld, float* ge0, float* ge1,
                         cudaStream_t s, bool infer, bool value_only = false) {
  FzArgs a{};
  fz_common(m, b, Wflat, wp, x, P, &a);
  const int top = b.top;
  for (int l = 0; l < top; ++l) fz_job(&a, b.wf_off[l], b.Np[l], b.Kp[l], l & 1, 1, (l & 1) + 1);
  fz_job(&a, b.wf_off[top], b.featN, b.Kp[top], top & 1, 1, 0);
  fz_job(&a, b.wtop_sdf_off, 16, b.Kp[top], (top & 1) ^ 1, 0, (top & 1) + 1);
  if (!value_only)
    for (int l = top - 1; l >= 0; --l) {
      const int st = top + 1 + (top - 1 - l);
      fz_job(&a, b.wt_off[l], l == 0 ? 64 : r16(m.in[l]), r64(m.out[l]), st & 1, 1, (st & 1) + 1);
    }
  a.sdf = sdf; a.sdf_ld = sdf_ld; a.has_feat = feat_b16 != nullptr; a.ge0 = ge0; a.ge1 = ge1; a.infer = infer;
  a.value_only = value_only;
  FzMaps maps{};
  const uint64_t LD = (uint64_t)b.LD, Pu = (uint64_t)P;
  if (int rc = make_tmap3(sv.pe, 64, Pu, 1, 64, Pu * 64, &maps.in0)) return rc;
  if (int rc = make_tmap3(sv.H, LD, Pu, (uint64_t)top, LD, Pu * LD, &maps.H)) return rc;
  if (infer || value_only) maps.D = maps.H;      // never stored: the saved block has no delta stack in inference / value-only mode
  else if (int rc = make_tmap3(sv.D, LD, Pu, (uint64_t)top, LD, Pu * LD, &maps.D)) return rc;
  if (feat_b16) {
    if (int rc = make_tmap3(feat_b16, (uint64_t)b.featN, Pu, 1, (uint64_t)feat_b16_ld, Pu * (uint64_t)feat_b16_ld, &maps.out)) return rc;
  } else {
    maps.out = maps.H;
  }
  maps.Z2 = maps.H;
  return launch_sdf_fused(FZ_FWD, a, maps, s);
}

int sdf_query_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, float* sdf_out, float* ws, cudaStream_t s,
                   bool ws_holds_pack, const bf16* wpx) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  if (P <= 0) return 0;
  bf16* wp_ws = reinterpret_cast<bf16*>(ws);
  const bf16* wp = wpx ? wpx : wp_ws;       // wpx: weights packed once per step behind the flat parameters (COPE_FLAT_HAS_PACK)
  bf16* pe = wp_ws + b.w_total;
  bf16* bufs[3] = {pe + P * 64, pe + P * 64 + P * b.LD, pe + P * 64 + 2 * P * b.LD};
  if (!ws_holds_pack && !wpx)   // ws_holds_pack: the caller vouches that an earlier query on this stream left the pack at the head of ws
    if (int rc = pack_sdf(m, b, Wflat, wp_ws, true, false, s)) return rc;
  if (sdf_chain_supported(m) && !getenv("COPE_NO_CHAIN")) {
    uint32_t offs[COPE_MAX_LIN];
    for (int l = 0; l < m.n_lin; ++l) offs[l] = (uint32_t)b.wf_off[l];
    return launch_sdf_chain_query(m, Wflat, wp, offs, (uint32_t)b.wtop_sdf_off, x, P, sdf_out, s);
  }
  bf16* hbuf[COPE_MAX_LIN + 1] = {nullptr};
  for (int l = 1; l <= b.top; ++l) hbuf[l] = (l == b.skip) ? bufs[2] : bufs[l & 1];
  pe_fwd_bf16_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, pe, 64, 64, b.skip > 0 ? hbuf[b.skip] + b.skw : nullptr, b.LD);
  COPE_CHECK_LAUNCH("pe_fwd_bf16");
  return sdf_layers_fwd(m, b, Wflat, wp, x, P, pe, hbuf, sdf_out, 1, nullptr, 0, s);
}

int sdf_fwd_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, float* sdf, int sdf_ld, float* feat,
                 int feat_ld, float* grad, float* saved, float* ws, cudaStream_t s, bf16* feat_b16, int feat_b16_ld, bool infer,
                 const bf16* wpx) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  if (P <= 0) return 0;
  SdfSavedB sv = sdf_saved_b(b, P, saved);
  bf16* wp_ws = reinterpret_cast<bf16*>(ws);
  const bf16* wp = wpx ? wpx : wp_ws;
  float* ge0 = reinterpret_cast<float*>(wp_ws + b.w_total);
  float* ge1 = ge0 + P * 64;
  if (!wpx)
    if (int rc = pack_sdf(m, b, Wflat, wp_ws, true, grad != nullptr, s)) return rc;
  if (!grad && !feat && !feat_b16 && sdf && fused_enabled(m, b, "fwd") && fused_enabled(m, b, "val")) {
    // value pass only (SDFNetwork.sdf with gradients, the SDF-consistency re-query of train.py:504): the chain stores H_1..H_top
    // for the value-only backward and stops before the reverse sweep
    return sdf_fwd_fused(m, b, Wflat, wp, x, P, sv, sdf, sdf_ld, nullptr, 0, ge0, ge1, s, false, true);
  }
  if (grad && !feat && fused_enabled(m, b, "fwd")) {
    if (int rc = sdf_fwd_fused(m, b, Wflat, wp, x, P, sv, sdf, sdf_ld, feat_b16, feat_b16_ld, ge0, ge1, s, infer)) return rc;
    pe_vjp_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, ge0, 64, b.skip > 0 ? ge1 : nullptr, 64, grad, m.d_in, 0);
    COPE_CHECK_LAUNCH("pe_vjp");
    return 0;
  }
  bf16* hbuf[COPE_MAX_LIN + 1] = {nullptr};
  for (int l = 1; l <= b.top; ++l) hbuf[l] = sv.h(l);
  pe_fwd_bf16_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, sv.pe, 64, 64, b.skip > 0 ? sv.h(b.skip) + b.skw : nullptr, b.LD);
  COPE_CHECK_LAUNCH("pe_fwd_bf16");
  if (int rc = sdf_layers_fwd(m, b, Wflat, wp, x, P, sv.pe, hbuf, sdf, sdf_ld, feat, feat_ld, s, feat_b16, feat_b16_ld)) return rc;
  if (!grad) return 0;
  // ---- reverse sweep
  const int top = b.top;
  bcast_sp_bf16_kernel<<<g1(P * (b.LD / 8)), 256, 0, s>>>(Wflat + m.w_off[top], sv.h(top), b.LD, hscale_of(b, top), sv.dl(top - 1),
                                                    b.LD, P, m.out[top - 1], b.LD);
  COPE_CHECK_LAUNCH("bcast_sp_bf16");
  if (b.skip > 0 && b.skw < b.LD) {
    zero_cols_bf16_kernel<<<g1(P * (b.LD - b.skw)), 256, 0, s>>>(sv.dl(b.skip - 1), b.LD, P, b.skw, b.LD);
    COPE_CHECK_LAUNCH("zero_cols");
  }
  for (int l = top - 1; l >= 0; --l) {
    const int nt = l == 0 ? 64 : r16(m.in[l]);
    const int kt = r64(m.out[l]);
    if (l > 0) {
      TcArgs t = tc_args((int)P, nt, kt, sv.dl(l), b.LD, wp + b.wt_off[l], sv.dl(l - 1), b.LD, 0);
      t.epi = TC_MUL_SIGP; t.H = sv.h(l); t.ldh = b.LD; t.hscale = hscale_of(b, l); t.alpha = alpha_of(b, l);
      t.n_valid = l == b.skip ? b.skw : m.in[l];
      if (l == b.skip) { t.nsplit = b.skw; t.out2 = ge1; t.ldo2 = 64; t.out2_f32 = 1; t.n2_valid = m.pe_w; }
      if (int rc = launch_tc_gemm(t, s)) return rc;
    } else {
      TcArgs t = tc_args((int)P, 64, kt, sv.dl(0), b.LD, wp + b.wt_off[0], ge0, 64, 1);
      t.n_valid = m.pe_w;
      if (int rc = launch_tc_gemm(t, s)) return rc;
    }
  }
  pe_vjp_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, ge0, 64, b.skip > 0 ? ge1 : nullptr, 64, grad, m.d_in, 0);
  COPE_CHECK_LAUNCH("pe_vjp");
  return 0;
}

// layer-by-layer tangent pass: T_{l+1} = alpha * (W_l T_l) * sp ; zb2_l = (W_l T_l) * delta_l * 100 (1 - sp)
static int sdf_tangent_layered(const MlpShape& m, const SdfB& b, const SdfSavedB& sv, const bf16* wp, const float* x, int64_t P,
                               const float* dgrad, bf16* T, bf16* t0, bf16* ZB2, float* dWflat, cudaStream_t s) {
  const int top = b.top, LD = b.LD;
  auto Tl = [&](int l) { return l == 0 ? t0 : T + (int64_t)(l - 1) * P * LD; };
  auto ldT = [&](int l) { return l == 0 ? 64 : LD; };
  auto zb2 = [&](int l) { return ZB2 + (int64_t)l * P * LD; };
  pe_jvp_bf16_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, dgrad, t0, 64, 64, b.skip > 0 ? Tl(b.skip) + b.skw : nullptr, LD);
  COPE_CHECK_LAUNCH("pe_jvp_bf16");
  for (int l = 0; l < top; ++l) {
    TcArgs t = tc_args((int)P, b.Np[l], b.Kp[l], Tl(l), ldT(l), wp + b.wf_off[l], Tl(l + 1), LD, 0);
    t.epi = TC_TANGENT; t.alpha = alpha_of(b, l + 1); t.H = sv.h(l + 1); t.ldh = LD; t.hscale = hscale_of(b, l + 1);
    t.D = sv.dl(l); t.ldd = LD; t.out2 = zb2(l); t.ldo2 = LD; t.out2_f32 = 0; t.n_valid = m.out[l];
    if (int rc = launch_tc_gemm(t, s)) return rc;
  }
  if (b.skip > 0 && b.skw < LD) {
    zero_cols_bf16_kernel<<<g1(P * (LD - b.skw)), 256, 0, s>>>(zb2(b.skip - 1), LD, P, b.skw, LD);
    COPE_CHECK_LAUNCH("zero_cols");
  }
  // last layer: delta_top = e_0  =>  dW_top[0, :] += sum_p T_top[p, :]
  return wcolsum(Tl(top), LD, nullptr, 0, P, m.in[top], dWflat + m.w_off[top], s);
}

// float offset of eb0 ([P x 64] fp32, followed by eb1) inside sdf_bwd_bf16's workspace (kernel-level tests)
int64_t sdf_bwd_eb_offset(const MlpShape& m, int64_t P) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  return ((int64_t)b.w_total + P * ((int64_t)(3 * b.top + 1) * b.LD + 64)) / 2;
}
// bf16 [P x LD] slot inside sdf_bwd_bf16's workspace that holds d_feat (written directly by the colour backward in the
// fused renderer path)
bf16* sdf_bwd_dfeat_slot(const MlpShape& m, int64_t P, float* ws, int* ld) {
  SdfB b;
  if (make_sdfb(m, &b)) return nullptr;
  bf16* wp = reinterpret_cast<bf16*>(ws);
  bf16* T = wp + b.w_total;
  bf16* ZB2 = T + (int64_t)b.top * P * b.LD;
  *ld = b.LD;
  return ZB2 + (int64_t)(2 * b.top) * P * b.LD;
}

// ------------------------------------------------------------------------------------------- SDF backward
int sdf_bwd_bf16(const MlpShape& m, const float* Wflat, const float* x, int64_t P, const float* saved, const float* d_sdf,
                 int d_sdf_ld, const float* d_feat, int d_feat_ld, const float* dgrad, float* dWflat, float* dx,
                 int dx_accumulate, float* ws, cudaStream_t s, bool d_feat_in_ws, const bf16* wpx) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  const bool have_dy = d_sdf || d_feat || d_feat_in_ws;
  if (P <= 0 || (!have_dy && !dgrad)) {
    if (dx && P > 0 && !dx_accumulate) cudaMemsetAsync(dx, 0, sizeof(float) * P * m.d_in, s);
    return 0;
  }
  SdfSavedB sv = sdf_saved_b(b, P, const_cast<float*>(saved));
  const int top = b.top, LD = b.LD;
  bf16* wp_ws = reinterpret_cast<bf16*>(ws);
  const bf16* wp = wpx ? wpx : wp_ws;
  bf16* T = wp_ws + b.w_total;                            // T_l, l = 1..top   -> T + (l-1) P LD
  bf16* ZB2 = T + (int64_t)top * P * LD;                  // zb2_l, l = 0..top-1
  bf16* ZBall = ZB2 + (int64_t)top * P * LD;              // zb_l, l = 0..top-1 (fused chain); the layer-by-layer path
  bf16* ZB[2] = {ZBall, ZBall + (int64_t)P * LD};         // ping-pongs between the first two
  bf16* dyb = ZBall + (int64_t)top * P * LD;              // bf16 copy of d_feat
  bf16* t0 = dyb + (int64_t)P * LD;                       // [P x 64]
  float* eb0 = reinterpret_cast<float*>(t0 + (int64_t)P * 64);
  float* eb1 = eb0 + P * 64;
  float* part = eb1 + P * 64;                            // tc_wgrad partial tiles
  auto Tl = [&](int l) { return l == 0 ? t0 : T + (int64_t)(l - 1) * P * LD; };
  auto ldT = [&](int l) { return l == 0 ? 64 : LD; };
  auto zb2 = [&](int l) { return ZB2 + (int64_t)l * P * LD; };
  if (!wpx)
    if (int rc = pack_sdf(m, b, Wflat, wp_ws, dgrad != nullptr, true, s)) return rc;
  const int featW = m.d_out - 1;

  const uint64_t LDu = (uint64_t)LD, Pu = (uint64_t)P;
  const bool fused_tan = dgrad && fused_enabled(m, b, "tan");
  if (fused_tan) {
    // ---- fused tangent sweep: T_1..T_top, zb2_0..zb2_{top-1} in one kernel
    FzMaps maps{};
    if (int rc = make_tmap3(sv.H, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.H)) return rc;
    if (int rc = make_tmap3(sv.D, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.D)) return rc;
    if (int rc = make_tmap3(ZB2, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.Z2)) return rc;
    FzArgs a{};
    fz_common(m, b, Wflat, wp, x, P, &a);
    a.g = dgrad;
    for (int l = 0; l < top; ++l) fz_job(&a, b.wf_off[l], b.Np[l], b.Kp[l], l & 1, 1, (l & 1) + 1);
    if (int rc = make_tmap3(t0, 64, Pu, 1, 64, Pu * 64, &maps.in0)) return rc;
    if (int rc = make_tmap3(T, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.out)) return rc;
    if (int rc = launch_sdf_fused(FZ_TAN, a, maps, s)) return rc;
    if (int rc = wcolsum(Tl(top), LD, nullptr, 0, P, m.in[top], dWflat + m.w_off[top], s)) return rc;
  }
  if (dgrad && (d_feat || d_feat_in_ws) && fused_enabled(m, b, "adj")) {
    // ---- fused adjoint sweep (stores every zb_l), then all weight gradients, then the value-path adjoint for dx
    FzMaps maps{};
    if (int rc = make_tmap3(sv.H, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.H)) return rc;
    if (int rc = make_tmap3(sv.D, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.D)) return rc;
    if (int rc = make_tmap3(ZB2, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.Z2)) return rc;
    if (!fused_tan) {
      if (int rc = sdf_tangent_layered(m, b, sv, wp, x, P, dgrad, T, t0, ZB2, dWflat, s)) return rc;
    }
    if (!d_feat_in_ws) {
      cvt_f32_bf16_kernel<<<g1(P * (r64(featW) / 8)), 256, 0, s>>>(d_feat, d_feat_ld, dyb, LD, P, featW, r64(featW), 1.0f);
      COPE_CHECK_LAUNCH("cvt_dfeat");
    }
    if (int rc = make_tmap3(dyb, LDu, Pu, 1, LDu, Pu * LDu, &maps.in0)) return rc;
    if (int rc = make_tmap3(ZBall, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.out)) return rc;
    for (int pass = 0; pass < (dx ? 2 : 1); ++pass) {
      FzArgs a{};
      fz_common(m, b, Wflat, wp, x, P, &a);
      a.d_sdf = d_sdf; a.d_sdf_ld = d_sdf_ld; a.eb0 = eb0; a.eb1 = eb1;
      a.has_d = pass == 0; a.store_out = pass == 0; a.want_e = pass == 1;
      fz_job(&a, b.wt_off[top], r16(m.in[top]), r64(featW), 0, 2, 1);
      for (int l = top - 1; l >= 1; --l) {
        const int st = top - l;
        fz_job(&a, b.wt_off[l], r16(m.in[l]), r64(m.out[l]), st & 1, 1, (st & 1) + 1);
      }
      if (a.want_e) fz_job(&a, b.wt_off[0], 64, r64(m.out[0]), top & 1, 1, (top & 1) + 1);
      if (int rc = launch_sdf_fused(FZ_ADJ, a, maps, s)) return rc;
      if (pass == 0) {
        // ---- weight / bias gradients from the stored adjoints: all layers in ONE launch
        TcWgradArgs wg[COPE_MAX_LIN];
        int nw = 0;
        {
          TcWgradArgs w{};
          w.P = P; w.Mp = r128(featW); w.Np = r16(m.in[top]); w.m_valid = featW; w.n_valid = m.in[top];
          w.X[0] = dyb; w.ldx[0] = LD; w.Y[0] = sv.h(top); w.ldy[0] = LD; w.n_pairs = 1;
          w.dW = dWflat + m.w_off[top] + m.in[top]; w.ldw = m.in[top]; w.part = part; w.db = dWflat + m.b_off[top] + 1;
          wg[nw++] = w;
          if (d_sdf) {
            if (int rc = wcolsum(sv.h(top), LD, d_sdf, d_sdf_ld, P, m.in[top], dWflat + m.w_off[top], s)) return rc;
            sum_strided_kernel<<<(unsigned)std::min<int64_t>(296, ceil_div(P, 256)), 256, 0, s>>>(d_sdf, d_sdf_ld, P, dWflat + m.b_off[top]);
            COPE_CHECK_LAUNCH("colsum_dsdf");
          }
        }
        for (int l = top - 1; l >= 0; --l) {
          TcWgradArgs w{};
          w.P = P; w.Mp = r128(m.out[l]); w.Np = l == 0 ? 64 : r16(m.in[l]); w.m_valid = m.out[l];
          w.n_valid = l == 0 ? m.pe_w : m.in[l];
          w.X[0] = ZBall + (int64_t)l * P * LD; w.ldx[0] = LD; w.Y[0] = sv.in(l); w.ldy[0] = sv.ld_in(l);
          w.X[1] = sv.dl(l); w.ldx[1] = LD; w.Y[1] = Tl(l); w.ldy[1] = ldT(l); w.n_pairs = 2;
          w.dW = dWflat + m.w_off[l]; w.ldw = m.in[l]; w.part = part; w.db = dWflat + m.b_off[l];
          wg[nw++] = w;
        }
        if (int rc = launch_tc_wgrad_batch(wg, nw, s)) return rc;
      } else {
        pe_vjp_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, eb0, 64, b.skip > 0 ? eb1 : nullptr, 64, dx, m.d_in, dx_accumulate);
        COPE_CHECK_LAUNCH("pe_vjp");
      }
    }
    return 0;
  }

  if (!dgrad && have_dy && fused_enabled(m, b, "adj") && fused_enabled(m, b, "val")) {
    // ---- value-only backward (SDFNetwork.forward / .sdf with gradients, e.g. the SDF-consistency loss of train.py:504): ONE
    // fused adjoint sweep (stores every zb_l, and eb0 / eb1 when dx is wanted), then all weight gradients in one launch
    if (d_feat) {
      cvt_f32_bf16_kernel<<<g1(P * (r64(featW) / 8)), 256, 0, s>>>(d_feat, d_feat_ld, dyb, LD, P, featW, r64(featW), 1.0f);
      COPE_CHECK_LAUNCH("cvt_dfeat");
    } else if (!d_feat_in_ws) {
      cudaMemsetAsync(dyb, 0, sizeof(bf16) * P * LD, s);
    }
    FzMaps maps{};
    if (int rc = make_tmap3(sv.H, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.H)) return rc;
    maps.D = maps.H; maps.Z2 = maps.H;
    if (int rc = make_tmap3(dyb, LDu, Pu, 1, LDu, Pu * LDu, &maps.in0)) return rc;
    if (int rc = make_tmap3(ZBall, LDu, Pu, (uint64_t)top, LDu, Pu * LDu, &maps.out)) return rc;
    FzArgs a{};
    fz_common(m, b, Wflat, wp, x, P, &a);
    a.d_sdf = d_sdf; a.d_sdf_ld = d_sdf_ld; a.eb0 = eb0; a.eb1 = eb1;
    a.has_d = 0; a.store_out = 1; a.want_e = dx != nullptr;
    fz_job(&a, b.wt_off[top], r16(m.in[top]), r64(featW), 0, 2, 1);
    for (int l = top - 1; l >= 1; --l) {
      const int st = top - l;
      fz_job(&a, b.wt_off[l], r16(m.in[l]), r64(m.out[l]), st & 1, 1, (st & 1) + 1);
    }
    if (a.want_e) fz_job(&a, b.wt_off[0], 64, r64(m.out[0]), top & 1, 1, (top & 1) + 1);
    if (int rc = launch_sdf_fused(FZ_ADJ, a, maps, s)) return rc;
    TcWgradArgs wg[COPE_MAX_LIN];
    int nw = 0;
    if (d_feat || d_feat_in_ws) {
      TcWgradArgs w{};
      w.P = P; w.Mp = r128(featW); w.Np = r16(m.in[top]); w.m_valid = featW; w.n_valid = m.in[top];
      w.X[0] = dyb; w.ldx[0] = LD; w.Y[0] = sv.h(top); w.ldy[0] = LD; w.n_pairs = 1;
      w.dW = dWflat + m.w_off[top] + m.in[top]; w.ldw = m.in[top]; w.part = part; w.db = dWflat + m.b_off[top] + 1;
      wg[nw++] = w;
    }
    if (d_sdf) {
      if (int rc = wcolsum(sv.h(top), LD, d_sdf, d_sdf_ld, P, m.in[top], dWflat + m.w_off[top], s)) return rc;
      sum_strided_kernel<<<(unsigned)std::min<int64_t>(296, ceil_div(P, 256)), 256, 0, s>>>(d_sdf, d_sdf_ld, P, dWflat + m.b_off[top]);
      COPE_CHECK_LAUNCH("colsum_dsdf");
    }
    for (int l = top - 1; l >= 0; --l) {
      TcWgradArgs w{};
      w.P = P; w.Mp = r128(m.out[l]); w.Np = l == 0 ? 64 : r16(m.in[l]); w.m_valid = m.out[l];
      w.n_valid = l == 0 ? m.pe_w : m.in[l];
      w.X[0] = ZBall + (int64_t)l * P * LD; w.ldx[0] = LD; w.Y[0] = sv.in(l); w.ldy[0] = sv.ld_in(l); w.n_pairs = 1;
      w.dW = dWflat + m.w_off[l]; w.ldw = m.in[l]; w.part = part; w.db = dWflat + m.b_off[l];
      wg[nw++] = w;
    }
    if (int rc = launch_tc_wgrad_batch(wg, nw, s)) return rc;
    if (dx) {
      pe_vjp_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, eb0, 64, b.skip > 0 ? eb1 : nullptr, 64, dx, m.d_in, dx_accumulate);
      COPE_CHECK_LAUNCH("pe_vjp");
    }
    return 0;
  }

  if (dgrad && !fused_tan) {
    if (int rc = sdf_tangent_layered(m, b, sv, wp, x, P, dgrad, T, t0, ZB2, dWflat, s)) return rc;
  }
  if (have_dy && !d_feat_in_ws) {
    cvt_f32_bf16_kernel<<<g1(P * (r64(featW) / 8)), 256, 0, s>>>(d_feat, d_feat_ld, dyb, LD, P, featW, r64(featW), 1.0f);
    COPE_CHECK_LAUNCH("cvt_dfeat");
  }

  const int n_pass = (dx && dgrad) ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    const bool do_w = pass == 0;
    const bool with2 = dgrad && pass == 0;
    const bool want_e = dx && (pass == n_pass - 1);
    if (!have_dy && !with2) break;
    const bf16* zb = nullptr;     // adjoint of z_l entering layer l's backward (l < top)
    int l_start = top - 1;
    if (have_dy) {
      // ---- top layer: zb_top = [d_sdf | d_feat]
      if (do_w) {
        TcWgradArgs w{};
        w.P = P; w.Mp = r128(featW); w.Np = r16(m.in[top]); w.m_valid = featW; w.n_valid = m.in[top];
        w.X[0] = dyb; w.ldx[0] = LD; w.Y[0] = sv.h(top); w.ldy[0] = LD; w.n_pairs = 1;
        w.dW = dWflat + m.w_off[top] + m.in[top]; w.ldw = m.in[top]; w.part = part;
        if (d_feat || d_feat_in_ws) {
          w.db = dWflat + m.b_off[top] + 1;
          if (int rc = launch_tc_wgrad(w, s)) return rc;
        }
        if (d_sdf) {
          if (int rc = wcolsum(sv.h(top), LD, d_sdf, d_sdf_ld, P, m.in[top], dWflat + m.w_off[top], s)) return rc;
          sum_strided_kernel<<<(unsigned)std::min<int64_t>(296, ceil_div(P, 256)), 256, 0, s>>>(d_sdf, d_sdf_ld, P, dWflat + m.b_off[top]);
          COPE_CHECK_LAUNCH("colsum_dsdf");
        }
      }
      TcArgs t = tc_args((int)P, r16(m.in[top]), r64(featW), dyb, LD, wp + b.wt_off[top], ZB[top & 1], LD, 0);
      t.epi = TC_BWD; t.H = sv.h(top); t.ldh = LD; t.hscale = hscale_of(b, top); t.alpha = alpha_of(b, top);
      t.n_valid = m.in[top];
      if (with2) { t.D = zb2(top - 1); t.ldd = LD; }
      if (d_sdf) { t.r1 = d_sdf; t.r1_ld = d_sdf_ld; t.r1w = Wflat + m.w_off[top]; }
      if (int rc = launch_tc_gemm(t, s)) return rc;
      zb = ZB[top & 1];
    } else {
      zb = zb2(top - 1);
    }
    for (int l = l_start; l >= 0; --l) {
      if (do_w) {
        TcWgradArgs w{};
        w.P = P; w.Mp = r128(m.out[l]); w.Np = l == 0 ? 64 : r16(m.in[l]); w.m_valid = m.out[l];
        w.n_valid = l == 0 ? m.pe_w : m.in[l];
        w.X[0] = zb; w.ldx[0] = LD; w.Y[0] = sv.in(l); w.ldy[0] = sv.ld_in(l); w.n_pairs = 1;
        if (with2) { w.X[1] = sv.dl(l); w.ldx[1] = LD; w.Y[1] = Tl(l); w.ldy[1] = ldT(l); w.n_pairs = 2; }
        w.dW = dWflat + m.w_off[l]; w.ldw = m.in[l]; w.part = part; w.db = dWflat + m.b_off[l];
        if (int rc = launch_tc_wgrad(w, s)) return rc;
      }
      if (l == 0 && !want_e) break;
      const int kt = r64(m.out[l]);
      if (l > 0) {
        bf16* nxt = ZB[l & 1];
        TcArgs t = tc_args((int)P, r16(m.in[l]), kt, zb, LD, wp + b.wt_off[l], nxt, LD, 0);
        t.epi = TC_BWD; t.H = sv.h(l); t.ldh = LD; t.hscale = hscale_of(b, l); t.alpha = alpha_of(b, l);
        t.n_valid = l == b.skip ? b.skw : m.in[l];
        if (with2) { t.D = zb2(l - 1); t.ldd = LD; }
        if (l == b.skip) {
          t.nsplit = b.skw; t.n2_valid = m.pe_w;
          if (want_e) { t.out2 = eb1; t.ldo2 = 64; t.out2_f32 = 1; }
        }
        if (int rc = launch_tc_gemm(t, s)) return rc;
        if (l == b.skip && b.skw < LD) {   // K padding of the next GEMM must be finite
          zero_cols_bf16_kernel<<<g1(P * (LD - b.skw)), 256, 0, s>>>(nxt, LD, P, b.skw, LD);
          COPE_CHECK_LAUNCH("zero_cols");
        }
        zb = nxt;
      } else {
        TcArgs t = tc_args((int)P, 64, kt, zb, LD, wp + b.wt_off[0], eb0, 64, 1);
        t.n_valid = m.pe_w;
        if (int rc = launch_tc_gemm(t, s)) return rc;
      }
    }
    if (want_e) {
      pe_vjp_kernel<<<g1(P * m.d_in), 256, 0, s>>>(x, P, m.d_in, m.L, eb0, 64, (b.skip > 0 && l_start >= b.skip) ? eb1 : nullptr,
                                                   64, dx, m.d_in, dx_accumulate);
      COPE_CHECK_LAUNCH("pe_vjp");
    }
  }
  if (dx && !have_dy && !dx_accumulate) cudaMemsetAsync(dx, 0, sizeof(float) * P * m.d_in, s);
  return 0;
}

// =========================================================================================== colour
// cin (bf16, ld CK = r64(in0 + 4)): [feat | x_hi(4) | PE(dirs) | normals(4) | x_lo(4) | 0]
struct ColB {
  int n_lin, top, LD, CK, d_feat, pe_w, rest, in0, d_out;
  size_t wf_off[COPE_MAX_LIN], wt_off[COPE_MAX_LIN], wt0_rest_off, w_total;
};
static int make_colb(const MlpShape& m, int Lv_or_neg, ColB* c) {
  c->n_lin = m.n_lin; c->top = m.n_lin - 1; c->in0 = m.in[0]; c->d_out = m.d_out;
  c->LD = 128;
  for (int l = 1; l < m.n_lin; ++l) {
    COPE_REQUIRE(m.in[l] % 64 == 0 && m.in[l] <= 256, "bf16 path: colour hidden width %d must be a multiple of 64, <= 256", m.in[l]);
    c->LD = std::max(c->LD, m.in[l]);
  }
  COPE_REQUIRE(m.d_out <= 16, "bf16 path: colour d_out=%d > 16", m.d_out);
  if (Lv_or_neg >= 0) {
    c->pe_w = 3 * (1 + 2 * Lv_or_neg);
    c->rest = 4 + c->pe_w + 4;
    c->d_feat = m.in[0] - c->rest;
    COPE_REQUIRE(c->d_feat > 0 && c->d_feat % 64 == 0 && c->d_feat <= 256 && c->rest + 4 <= 64,
                 "bf16 path: colour input %d does not split into feat(%%64) + <=60 extras", m.in[0]);
  } else {                       // size queries do not know Lv: bound it
    c->pe_w = 27; c->rest = 35; c->d_feat = m.in[0] - 35;
    if (c->d_feat <= 0) c->d_feat = 64;
  }
  c->CK = r64(c->d_feat) + 64;
  size_t off = 0;
  for (int l = 0; l < m.n_lin; ++l) {
    const int kp = l == 0 ? c->CK : r64(m.in[l]);
    c->wf_off[l] = off; off += (size_t)r16(m.out[l]) * kp;
    const int nt = l == 0 ? r16(c->d_feat) : r16(m.in[l]);
    c->wt_off[l] = off; off += (size_t)nt * r64(m.out[l]);
  }
  c->wt0_rest_off = off; off += (size_t)64 * r64(m.out[0]);
  c->w_total = (off + 63) / 64 * 64;
  return 0;
}

static int pack_color(const MlpShape& m, const ColB& c, const float* Wflat, bf16* wp, bool fwd, bool transposed, cudaStream_t s) {
  PackBatch pb{};
  const int F = c.d_feat, R = c.rest;
  for (int l = 0; l < m.n_lin; ++l) {
    const float* W = Wflat + m.w_off[l];
    if (fwd) {
      PackSpec sp = pack_spec();
      seg_n(sp, 0, 0, m.out[l]);
      if (l == 0) { seg_k(sp, 0, R, F); seg_k(sp, F, 0, R); seg_k(sp, F + R, 0, 4); } else seg_k(sp, 0, 0, m.in[l]);
      pack_add(pb, W, m.in[l], sp, r16(m.out[l]), l == 0 ? c.CK : r64(m.in[l]), 0, wp + c.wf_off[l]);
    }
    if (transposed) {
      PackSpec sp = pack_spec();
      if (l == 0) seg_n(sp, 0, R, F); else seg_n(sp, 0, 0, m.in[l]);
      seg_k(sp, 0, 0, m.out[l]);
      pack_add(pb, W, m.in[l], sp, l == 0 ? r16(F) : r16(m.in[l]), r64(m.out[l]), 1, wp + c.wt_off[l]);
      if (l == 0) {
        PackSpec sr = pack_spec();
        seg_n(sr, 0, 0, R); seg_k(sr, 0, 0, m.out[0]);
        pack_add(pb, W, m.in[0], sr, 64, r64(m.out[0]), 1, wp + c.wt0_rest_off);
      }
    }
  }
  return launch_tc_pack_batch(pb, s);
}

__device__ __forceinline__ float color_in_elem(const float* x, const float* dv, int Lv, const float* nrm, int e, int pe_w, int R) {
  // e indexes the non-feature tail: [x_hi(4) | PE(dirs) | normals(4) | x_lo(4) | 0...]
  if (e < 4) return __bfloat162float(__float2bfloat16(x[e]));
  if (e < 4 + pe_w) {
    const int q = e - 4;
    if (q < 3) return dv[q];
    const int blk = (q - 3) / 3, dd = (q - 3) % 3;
    const float a = dv[dd] * (float)(1 << (blk >> 1));
    return (blk & 1) ? cosf(a) : sinf(a);
  }
  if (e < R) return nrm[e - 4 - pe_w];
  if (e < R + 4) { const float xv = x[e - R]; return xv - __bfloat162float(__float2bfloat16(xv)); }
  return 0.0f;
}
// one thread = 8 consecutive columns of one row (16-byte stores; feature columns are 32-byte fp32 reads)
__global__ void color_pack_bf16_kernel(const float* __restrict__ x, const float* __restrict__ dirs, int dirs_group, int Lv,
                                       const float* __restrict__ nrm, const float* __restrict__ feat, int feat_ld, int F,
                                       int64_t P, bf16* __restrict__ cin, int ld, int col0) {
  // col0 > 0: the first col0 (= F) columns already hold the bf16 features; only the tail is written
  const int pe_w = 3 * (1 + 2 * Lv);
  const int R = 4 + pe_w + 4;
  const int groups = (ld - col0) >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * groups) return;
  const int64_t p = i / groups;
  const int c0 = col0 + (int)(i - p * groups) * 8;
  float v[8];
  if (c0 + 8 <= F) {
    const float* f = feat + p * feat_ld + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = f[k];
  } else {
    const float* dv = dirs + (p / dirs_group) * 3;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      v[k] = c < F ? feat[p * feat_ld + c] : color_in_elem(x + p * 4, dv, Lv, nrm + p * 4, c - F, pe_w, R);
    }
  }
  uint4 q;
  q.x = tc::pack_bf16(v[0], v[1]); q.y = tc::pack_bf16(v[2], v[3]); q.z = tc::pack_bf16(v[4], v[5]); q.w = tc::pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(cin + p * ld + c0) = q;
}

// dz_top[p, c] = d_rgb * rgb (1 - rgb)  (bf16, zero padded to `ld`)
__global__ void sigmoid_bwd_bf16_kernel(const float* __restrict__ dout, const float* __restrict__ out, int n, bf16* __restrict__ dz,
                                        int ld, int64_t P) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * ld) return;
  int64_t p = i / ld;
  int c = (int)(i - p * ld);
  float v = 0.0f;
  if (c < n) { float o = out[p * n + c]; v = dout[p * n + c] * o * (1.0f - o); }
  dz[i] = __float2bfloat16(v);
}

// rest[p, :] (fp32, ld 64) = [dx(4) | dPE(dirs) | dnormals(4)]
__global__ void color_unpack_rest_kernel(const float* __restrict__ rest, const float* __restrict__ dirs, int dirs_group, int Lv,
                                         int64_t P, float* __restrict__ dx, float* __restrict__ ddirs, float* __restrict__ dnrm) {
  const int pe_w = 3 * (1 + 2 * Lv);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * 11) return;
  int64_t p = i / 11;
  int c = (int)(i - p * 11);
  const float* r = rest + p * 64;
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
  } else if (dnrm) dnrm[p * 4 + (c - 7)] += r[4 + pe_w + (c - 7)];
}

struct ColSavedB {
  bf16* cin; bf16* H; float* rgb; int64_t P; int LD, CK;
  bf16* h(int l) const { return H + (int64_t)(l - 1) * P * LD; }
  const bf16* in(int l) const { return l == 0 ? cin : h(l); }
  int ld_in(int l) const { return l == 0 ? CK : LD; }
};
static ColSavedB col_saved_b(const ColB& c, int64_t P, float* base) {
  ColSavedB v; v.P = P; v.LD = c.LD; v.CK = c.CK;
  v.rgb = base;                                               // fp32 [P x d_out] first (alignment)
  v.cin = reinterpret_cast<bf16*>(base + ((P * c.d_out + 63) / 64 * 64));
  v.H = v.cin + P * c.CK;
  return v;
}
bf16* color_cin_slot(const MlpShape& m, int Lv, int64_t P, float* saved, int* ld) {
  ColB c;
  if (make_colb(m, Lv, &c)) return nullptr;
  *ld = c.CK;
  return col_saved_b(c, P, saved).cin;
}
int64_t color_saved_floats_bf16(const MlpShape& m, int64_t P) {
  ColB c;
  if (make_colb(m, -1, &c)) return -1;
  return (P * c.d_out + 63) / 64 * 64 + (P * (c.CK + (int64_t)c.top * c.LD) + 1) / 2 + 64;
}
int64_t color_ws_floats_bf16(const MlpShape& m, int64_t P) {
  ColB c;
  if (make_colb(m, -1, &c)) return -1;
  return ((int64_t)c.w_total + P * ((int64_t)c.top * c.LD + 128) + 1) / 2 + P * 64 + tc_wgrad_part_floats() + 1024;
}

static bool color_fused_enabled(const MlpShape& m, const ColB& c) {
  const char* e = getenv("COPE_NO_FUSED");
  if (e && (strstr(e, "col") || strstr(e, "all") || !strcmp(e, "1"))) return false;
  return c.LD == 256 && c.CK == 320 && color_fused_supported(m, c.d_feat, c.rest);
}
static void cz_job(CzArgs* a, size_t w_off, int Np, int Kp, int acc, int wait_a, int commit, int drain) {
  FzJob& j = a->jobs[a->n_jobs++];
  j.w_off = (uint32_t)w_off; j.Np = (uint16_t)Np; j.Kp = (uint16_t)Kp; j.acc = (uint8_t)acc; j.wait_a = (uint8_t)wait_a;
  j.commit = (uint8_t)commit; j.pad = (uint8_t)drain;
}

int color_fwd_bf16(const MlpShape& m, const float* Wflat, const float* x, const float* dirs, int dirs_group, int Lv,
                   const float* normals, const float* feat, int feat_ld, int64_t P, float* rgb, float* saved, float* ws,
                   cudaStream_t s, bool feat_in_cin, bool infer, const bf16* wpx) {
  ColB c;
  if (make_colb(m, Lv, &c)) return -1;
  COPE_REQUIRE(m.skip < 0, "bf16 colour path: skip connections are not supported");
  if (P <= 0) return 0;
  ColSavedB sv = col_saved_b(c, P, saved);
  bf16* wp_ws = reinterpret_cast<bf16*>(ws);
  const bf16* wp = wpx ? wpx : wp_ws;
  if (!wpx)
    if (int rc = pack_color(m, c, Wflat, wp_ws, true, false, s)) return rc;
  if (feat_in_cin && color_fused_enabled(m, c)) {
    // ---- fused chain: tail of the input built in the kernel, 4 ReLU layers + sigmoid, h_l TMA-stored for the backward
    CzArgs a{};
    a.P = P; a.Wflat = Wflat; a.wp = wp; a.n_lin = m.n_lin; a.d_out = m.d_out;
    for (int l = 0; l < m.n_lin; ++l) a.b_off[l] = m.b_off[l];
    a.x = x; a.dirs = dirs; a.dirs_group = dirs_group; a.Lv = Lv; a.normals = normals; a.rgb = rgb; a.rgb_saved = infer ? nullptr : sv.rgb;
    a.infer = infer;
    cz_job(&a, c.wf_off[0], r16(m.out[0]), c.CK, 0, 3, 1, 0);
    for (int l = 1; l < c.top; ++l) cz_job(&a, c.wf_off[l], r16(m.out[l]), r64(m.in[l]), l & 1, 1, (l & 1) + 1, 0);
    cz_job(&a, c.wf_off[c.top], r16(m.out[c.top]), r64(m.in[c.top]), c.top & 1, 1, (c.top & 1) + 1, 0);
    CzMaps maps{};
    const uint64_t Pu = (uint64_t)P, CK = (uint64_t)c.CK, LD = (uint64_t)c.LD;
    if (int rc = make_tmap3(sv.cin, (uint64_t)c.d_feat, Pu, 1, CK, Pu * CK, &maps.feat)) return rc;
    if (int rc = make_tmap3(sv.cin + c.d_feat, 64, Pu, 1, CK, Pu * CK, &maps.tail)) return rc;
    if (int rc = make_tmap3(sv.H, LD, Pu, (uint64_t)c.top, LD, Pu * LD, &maps.H)) return rc;
    maps.DZ = maps.H;
    return launch_color_fused(CZ_FWD, a, maps, s);
  }
  color_pack_bf16_kernel<<<g1(P * ((feat_in_cin ? c.CK - c.d_feat : c.CK) / 8)), 256, 0, s>>>(
      x, dirs, dirs_group, Lv, normals, feat, feat_ld, c.d_feat, P, sv.cin, c.CK, feat_in_cin ? c.d_feat : 0);
  COPE_CHECK_LAUNCH("color_pack_bf16");
  for (int l = 0; l < m.n_lin; ++l) {
    const bool last = l == c.top;
    TcArgs t = tc_args((int)P, r16(m.out[l]), l == 0 ? c.CK : r64(m.in[l]), sv.in(l), sv.ld_in(l), wp + c.wf_off[l],
                       last ? (void*)sv.rgb : (void*)sv.h(l + 1), last ? m.d_out : c.LD, last ? 1 : 0);
    t.bias = Wflat + m.b_off[l]; t.epi = last ? TC_BIAS_SIGMOID : TC_BIAS_RELU; t.n_valid = m.out[l];
    if (int rc = launch_tc_gemm(t, s)) return rc;
  }
  cudaMemcpyAsync(rgb, sv.rgb, sizeof(float) * P * m.d_out, cudaMemcpyDeviceToDevice, s);
  return 0;
}

int color_bwd_bf16(const MlpShape& m, const float* Wflat, const float* dirs, int dirs_group, int Lv, int64_t P,
                   const float* saved, const float* d_rgb, float* dWflat, float* dx, float* ddirs, float* dnormals,
                   float* dfeat, int dfeat_ld, float* ws, cudaStream_t s, bf16* dfeat_b16, int dfeat_b16_ld, const bf16* wpx) {
  ColB c;
  if (make_colb(m, Lv, &c)) return -1;
  if (P <= 0) return 0;
  ColSavedB sv = col_saved_b(c, P, const_cast<float*>(saved));
  bf16* wp_ws = reinterpret_cast<bf16*>(ws);
  const bf16* wp = wpx ? wpx : wp_ws;
  bf16* DZ = wp_ws + c.w_total;                                // dz_l, l = 0..top-1 (fused chain); the layer-by-layer
  bf16* B[2] = {DZ, DZ + (int64_t)P * c.LD};                   // path ping-pongs between the first two
  bf16* dzt = DZ + (int64_t)c.top * P * c.LD;                  // [P x 128]
  float* rest = reinterpret_cast<float*>(dzt + (int64_t)P * 128);
  float* part = rest + P * 64;
  if (!wpx)
    if (int rc = pack_color(m, c, Wflat, wp_ws, false, true, s)) return rc;
  if (!dfeat && color_fused_enabled(m, c)) {
    // ---- fused adjoint chain (stores every dz_l), then the weight gradients from the stored tiles
    const int F = c.d_feat, R = c.rest, top = c.top;
    const bool want_rest = dx || ddirs || dnormals;
    CzArgs a{};
    a.P = P; a.Wflat = Wflat; a.wp = wp; a.n_lin = m.n_lin; a.d_out = m.d_out;
    a.d_rgb = d_rgb; a.rgb_in = sv.rgb; a.rest = want_rest ? rest : nullptr; a.want_dfeat = dfeat_b16 != nullptr;
    cz_job(&a, c.wt_off[top], r16(m.in[top]), r64(m.out[top]), 0, 1, 1, 2);
    for (int l = top - 1; l >= 1; --l) {
      const int st = top - l;
      cz_job(&a, c.wt_off[l], r16(m.in[l]), r64(m.out[l]), st & 1, 1, (st & 1) + 1, 0);
    }
    cz_job(&a, c.wt0_rest_off, 64, r64(m.out[0]), 1, 1, 0, 0);
    cz_job(&a, c.wt_off[0], r16(F), r64(m.out[0]), 0, 0, 1, 0xF);
    CzMaps maps{};
    const uint64_t Pu = (uint64_t)P, LD = (uint64_t)c.LD;
    if (int rc = make_tmap3(dzt, 128, Pu, 1, 128, Pu * 128, &maps.tail)) return rc;
    if (int rc = make_tmap3(sv.H, LD, Pu, (uint64_t)top, LD, Pu * LD, &maps.H)) return rc;
    if (int rc = make_tmap3(DZ, LD, Pu, (uint64_t)top, LD, Pu * LD, &maps.DZ)) return rc;
    if (dfeat_b16) {
      if (int rc = make_tmap3(dfeat_b16, (uint64_t)r16(F), Pu, 1, (uint64_t)dfeat_b16_ld, Pu * (uint64_t)dfeat_b16_ld, &maps.feat)) return rc;
    } else {
      maps.feat = maps.DZ;
    }
    if (int rc = launch_color_fused(CZ_BWD, a, maps, s)) return rc;
    TcWgradArgs wg[COPE_MAX_LIN + 1];
    int nw = 0;
    for (int l = top; l >= 0; --l) {
      TcWgradArgs w{};
      w.P = P; w.Mp = r128(m.out[l]); w.m_valid = m.out[l]; w.n_pairs = 1; w.part = part;
      w.X[0] = l == top ? dzt : DZ + (int64_t)l * P * c.LD; w.ldx[0] = l == top ? 128 : c.LD;
      w.db = dWflat + m.b_off[l];
      if (l > 0) {
        w.Np = r16(m.in[l]); w.n_valid = m.in[l]; w.Y[0] = sv.h(l); w.ldy[0] = c.LD; w.dW = dWflat + m.w_off[l]; w.ldw = m.in[l];
        wg[nw++] = w;
      } else {
        w.Np = r16(F); w.n_valid = F; w.Y[0] = sv.cin; w.ldy[0] = c.CK; w.dW = dWflat + m.w_off[0] + R; w.ldw = m.in[0];
        wg[nw++] = w;
        w.db = nullptr;
        w.Np = 64; w.n_valid = R; w.Y[0] = sv.cin + r64(F); w.dW = dWflat + m.w_off[0];
        wg[nw++] = w;
      }
    }
    if (int rc = launch_tc_wgrad_batch(wg, nw, s)) return rc;
    if (want_rest) {
      color_unpack_rest_kernel<<<g1(P * 11), 256, 0, s>>>(rest, dirs, dirs_group, Lv, P, dx, ddirs, dnormals);
      COPE_CHECK_LAUNCH("color_unpack_rest");
    }
    return 0;
  }
  sigmoid_bwd_bf16_kernel<<<g1(P * 128), 256, 0, s>>>(d_rgb, sv.rgb, m.d_out, dzt, 128, P);
  COPE_CHECK_LAUNCH("sigmoid_bwd_bf16");
  const bf16* dz = dzt;
  int lddz = 128;
  const int F = c.d_feat, R = c.rest;
  for (int l = c.top; l >= 0; --l) {
    // ---- weight + bias gradients
    TcWgradArgs w{};
    w.P = P; w.Mp = r128(m.out[l]); w.m_valid = m.out[l]; w.X[0] = dz; w.ldx[0] = lddz; w.n_pairs = 1; w.part = part;
    w.db = dWflat + m.b_off[l];                  // bias gradient: fused column sum of dz inside the wgrad kernel
    if (l > 0) {
      w.Np = r16(m.in[l]); w.n_valid = m.in[l]; w.Y[0] = sv.h(l); w.ldy[0] = c.LD; w.dW = dWflat + m.w_off[l]; w.ldw = m.in[l];
      if (int rc = launch_tc_wgrad(w, s)) return rc;
    } else {
      w.Np = r16(F); w.n_valid = F; w.Y[0] = sv.cin; w.ldy[0] = c.CK; w.dW = dWflat + m.w_off[0] + R; w.ldw = m.in[0];
      if (int rc = launch_tc_wgrad(w, s)) return rc;
      w.db = nullptr;
      w.Np = 64; w.n_valid = R; w.Y[0] = sv.cin + r64(F); w.dW = dWflat + m.w_off[0];
      if (int rc = launch_tc_wgrad(w, s)) return rc;
    }
    // ---- data gradient
    const int kt = r64(m.out[l]);
    if (l > 0) {
      bf16* nxt = B[l & 1];
      TcArgs t = tc_args((int)P, r16(m.in[l]), kt, dz, lddz, wp + c.wt_off[l], nxt, c.LD, 0);
      t.epi = TC_RELU_MASK; t.H = sv.h(l); t.ldh = c.LD; t.n_valid = m.in[l];
      if (int rc = launch_tc_gemm(t, s)) return rc;
      dz = nxt; lddz = c.LD;
    } else {
      if (dfeat || dfeat_b16) {
        TcArgs t = dfeat_b16 ? tc_args((int)P, r16(F), kt, dz, lddz, wp + c.wt_off[0], dfeat_b16, dfeat_b16_ld, 0)
                             : tc_args((int)P, r16(F), kt, dz, lddz, wp + c.wt_off[0], dfeat, dfeat_ld, 1);
        t.n_valid = F;
        if (int rc = launch_tc_gemm(t, s)) return rc;
      }
      if (dx || ddirs || dnormals) {
        TcArgs t = tc_args((int)P, 64, kt, dz, lddz, wp + c.wt0_rest_off, rest, 64, 1);
        t.n_valid = R;
        if (int rc = launch_tc_gemm(t, s)) return rc;
        color_unpack_rest_kernel<<<g1(P * 11), 256, 0, s>>>(rest, dirs, dirs_group, Lv, P, dx, ddirs, dnormals);
        COPE_CHECK_LAUNCH("color_unpack_rest");
      }
    }
  }
  return 0;
}

// workspace of the FORWARD entry points alone (inference render): sdf_fwd_bf16 uses [packed weights | ge0 | ge1], color_fwd_bf16
// only the packed weights - a quarter of the training workspace, which also holds the T / zb2 / zb stacks of the backward
int64_t sdf_fwd_ws_floats_bf16(const MlpShape& m, int64_t P) {
  SdfB b;
  if (make_sdfb(m, &b)) return -1;
  return ((int64_t)b.w_total + 1) / 2 + 2 * P * 64 + 64;
}
int64_t color_fwd_ws_floats_bf16(const MlpShape& m) {
  ColB c;
  if (make_colb(m, -1, &c)) return -1;
  return ((int64_t)c.w_total + 1) / 2 + 64;
}

// ------------------------------------------------------------------------------------------- weights packed once per step
// The packed bf16 operands of a network (forward AND transposed blocks, in the layout every bf16 entry point expects at `wp`)
// written behind its flat fp32 parameters: the five per-call re-packs of a training step become one launch per network.
int64_t mlp_pack_elems_bf16(const MlpShape& m, int is_color, int Lv) {
  if (is_color) { ColB c; if (make_colb(m, Lv, &c)) return -1; return (int64_t)c.w_total; }
  SdfB b; if (make_sdfb(m, &b)) return -1; return (int64_t)b.w_total;
}
int mlp_pack_bf16(const MlpShape& m, int is_color, int Lv, const float* Wflat, bf16* wp, cudaStream_t s) {
  if (is_color) { ColB c; if (make_colb(m, Lv, &c)) return -1; return pack_color(m, c, Wflat, wp, true, true, s); }
  SdfB b; if (make_sdfb(m, &b)) return -1; return pack_sdf(m, b, Wflat, wp, true, true, s);
}

}  // namespace cope
