// NeuSRenderer's non-MLP work as one-warp-per-ray kernels: coarse depths, ray points, up_sample
// (sigmoid-CDF alpha -> warp product scan -> pdf/cdf -> inverse-CDF search), sorted merge (cat_z_vals) and
// render_core's compositing forward/backward.  All HBM-bound: each sample is read once, lanes own contiguous
// chunks of the ray so loads coalesce, scans are warp shuffles.
//
// Arithmetic that feeds integer decisions (search indices, merge order) uses explicit _rn intrinsics so that
// nvcc does not contract a*b+c into an FMA the reference's ATen kernels do not use.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace cope {

constexpr int kWarpsPerBlock = 4;

// torch.linspace(start, end, steps)[i] in fp32 (ATen: symmetric evaluation around the midpoint)
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
  return i < steps / 2 ? __fadd_rn(start, __fmul_rn(step, (float)i))
                       : __fsub_rn(end, __fmul_rn(step, (float)(steps - i - 1)));
}

// ------------------------------------------------------------------------------------ coarse z / points
// neus_renderer.py:466-483
__global__ void coarse_z_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                const float* __restrict__ t_rand, int64_t N, int S, float* __restrict__ z) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * S) return;
  int64_t n = i / S;
  int j = (int)(i - n * S);
  float nr = near[n], fr = far[n];
  auto zc = [&](int q) {
    float u = linspace_at(0.0f, 1.0f, S, q);
    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, u)), __fmul_rn(fr, u));
  };
  float zj = zc(j);
  if (t_rand) {
    float lo = j > 0 ? __fmul_rn(0.5f, __fadd_rn(zj, zc(j - 1))) : zj;
    float hi = j < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zc(j + 1), zj)) : zj;
    zj = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[i]));
  }
  z[i] = zj;
}

// neus_renderer.py:337-350 (use_mid) and :495-498 / :285 (at z)
__global__ void ray_points_kernel(const float* __restrict__ ro, const float* __restrict__ rd,
                                  const float* __restrict__ z, const float* __restrict__ tstep,
                                  const float* __restrict__ near, const float* __restrict__ far, int n_coarse,
                                  int64_t N, int S, int use_mid, float4* __restrict__ pts, float* __restrict__ dists,
                                  float* __restrict__ mid_z) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * S) return;
  int64_t n = i / S;
  int j = (int)(i - n * S);
  float zj = z[i], zz = zj;
  if (use_mid) {
    float sd = __fdiv_rn(__fsub_rn(far[0], near[0]), (float)n_coarse);
    float dj = j < S - 1 ? __fsub_rn(z[i + 1], zj) : sd;
    zz = __fadd_rn(zj, __fmul_rn(dj, 0.5f));
    if (dists) dists[i] = dj;
    if (mid_z) mid_z[i] = zz;
  }
  const float* o = ro + n * 3;
  const float* d = rd + n * 3;
  float4 p;
  p.x = __fadd_rn(o[0], __fmul_rn(d[0], zz));
  p.y = __fadd_rn(o[1], __fmul_rn(d[1], zz));
  p.z = __fadd_rn(o[2], __fmul_rn(d[2], zz));
  p.w = tstep[0];
  pts[i] = p;
}

// d_o = sum_s d_pts ; d_d += sum_s d_pts * zz      (one warp per ray)
__global__ void ray_points_bwd_kernel(const float4* __restrict__ dpts, const float* __restrict__ zz,
                                      const float* __restrict__ ddirs, int64_t N, int S, float* __restrict__ d_o,
                                      float* __restrict__ d_d) {
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= N) return;
  int lane = threadIdx.x & 31;
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
  for (int j = lane; j < S; j += 32) {
    float4 g = dpts[n * S + j];
    float t = zz[n * S + j];
    ox += g.x; oy += g.y; oz += g.z;
    dx += g.x * t; dy += g.y * t; dz += g.z * t;
    if (ddirs) {
      const float* q = ddirs + (n * S + j) * 3;
      dx += q[0]; dy += q[1]; dz += q[2];
    }
  }
  ox = warp_sum(ox); oy = warp_sum(oy); oz = warp_sum(oz);
  dx = warp_sum(dx); dy = warp_sum(dy); dz = warp_sum(dz);
  if (lane == 0) {
    d_o[n * 3 + 0] = ox; d_o[n * 3 + 1] = oy; d_o[n * 3 + 2] = oz;
    d_d[n * 3 + 0] += dx; d_d[n * 3 + 1] += dy; d_d[n * 3 + 2] += dz;
  }
}

// ------------------------------------------------------------------------------------ inverse CDF
// neus_renderer.py:47-70 with det=True.  cdf/bins: S entries in shared memory.
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int S, int K, int k, int64_t* ind_out) {
  float u = linspace_at(0.5f / (float)K, 1.0f - 0.5f / (float)K, K, k);
  int lo = 0, hi = S;                       // searchsorted(right=True): first i with cdf[i] > u
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  if (ind_out) *ind_out = lo;
  int below = lo - 1 > 0 ? lo - 1 : 0;
  int above = lo < S - 1 ? lo : S - 1;
  float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
  float den = __fsub_rn(c1, c0);
  if (den < 1e-5f) den = 1.0f;
  float t = __fdiv_rn(__fsub_rn(u, c0), den);
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

constexpr int kMaxS = 256;   // samples per ray the sampling kernels hold in shared memory

__global__ void sample_cdf_kernel(const float* __restrict__ cdf, const float* __restrict__ bins, int64_t N, int S,
                                  int K, float* __restrict__ samples, int64_t* __restrict__ inds) {
  __shared__ float s_cdf[kWarpsPerBlock][kMaxS], s_bin[kWarpsPerBlock][kMaxS];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + w;
  if (n >= N) return;
  for (int j = lane; j < S; j += 32) { s_cdf[w][j] = cdf[n * S + j]; s_bin[w][j] = bins[n * S + j]; }
  __syncwarp();
  for (int k = lane; k < K; k += 32) {
    int64_t ind;
    samples[n * K + k] = invert_cdf(s_cdf[w], s_bin[w], S, K, k, &ind);
    if (inds) inds[n * K + k] = ind;
  }
}

// ------------------------------------------------------------------------------------ up_sample
// neus_renderer.py:178-224 + sample_pdf :39-70.  Lane owns sections [lane*C, lane*C+C).
template <int MAXC>      // sections per lane the loops are unrolled for: 4 (S <= 129) or 8 (S <= kMaxS)
__global__ void upsample_kernel(const float* __restrict__ z, const float* __restrict__ sdf, int64_t N, int S, int K,
                                float inv_s, float* __restrict__ new_z, float* __restrict__ cdf_out,
                                int64_t* __restrict__ inds_out) {
  __shared__ float s_z[kWarpsPerBlock][kMaxS], s_f[kWarpsPerBlock][kMaxS], s_cdf[kWarpsPerBlock][kMaxS];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + w;
  if (n >= N) return;
  {
    // all loads of the row in flight together (S <= 32 * (MAXC + 1); the last pass covers S - 1 = 32 * MAXC)
    float zr[MAXC + 1], fr[MAXC + 1];
#pragma unroll
    for (int c = 0; c <= MAXC; ++c) {
      const int j = lane + 32 * c;
      zr[c] = j < S ? z[n * S + j] : 0.0f;
      fr[c] = j < S ? sdf[n * S + j] : 0.0f;
    }
#pragma unroll
    for (int c = 0; c <= MAXC; ++c) {
      const int j = lane + 32 * c;
      if (j < S) { s_z[w][j] = zr[c]; s_f[w][j] = fr[c]; }
    }
  }
  __syncwarp();
  const float* Z = s_z[w];
  const float* F = s_f[w];
  const int M = S - 1;                      // sections
  const int C = (M + 31) / 32;
  float alpha[MAXC], wgt[MAXC];
  // The kernel is instruction-bound (one warp turns ~1 KB of a ray into 16 samples), so the per-section arithmetic uses the
  // fast division / exp intrinsics: |error| of alpha ~3e-7, of the CDF ~1e-6 (tests: 2e-6), far below what moves a sample by
  // more than the 2e-5 the parity test allows.  The inverse-CDF step (invert_cdf) keeps the IEEE operations: it is the part
  // that is bit-exact against torch.searchsorted when fed the reference's CDF.  raw_cos(j-1) is carried from the previous
  // section of the lane instead of being recomputed.
  auto raw_cos = [&](int j) { return __fdividef(F[j + 1] - F[j], (Z[j + 1] - Z[j]) + 1e-5f); };
  auto fast_sigmoid = [](float x) { return __fdividef(1.0f, 1.0f + __expf(fminf(-x, 80.0f))); };
  float prod = 1.0f;
  float pv = 0.0f;
  {
    const int j0 = lane * C;
    if (j0 > 0 && j0 < M) pv = raw_cos(j0 - 1);
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    alpha[c] = 0.0f;
    if (c < C && j < M) {
      const float raw = raw_cos(j);
      const float cs = fminf(fmaxf(fminf(pv, raw), -1e3f), 0.0f);
      pv = raw;
      const float mid = (F[j] + F[j + 1]) * 0.5f;
      const float dist = Z[j + 1] - Z[j];
      const float half = cs * dist * 0.5f;
      const float pc = fast_sigmoid((mid - half) * inv_s);
      const float nc = fast_sigmoid((mid + half) * inv_s);
      alpha[c] = __fdividef(pc - nc + 1e-5f, pc + 1e-5f);
      prod *= 1.0f - alpha[c] + 1e-7f;
    }
  }
  // exclusive product scan of the per-lane products
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= v;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.0f;
  float lsum = 0.0f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    wgt[c] = 0.0f;
    if (c < C && j < M) {
      wgt[c] = alpha[c] * T + 1e-5f;                          // weights + 1e-5 (:42)
      T *= 1.0f - alpha[c] + 1e-7f;
      lsum += wgt[c];
    }
  }
  float total = warp_sum(lsum);
  // inclusive sum scan of pdf
  float lp = 0.0f;
  const float inv_total = __frcp_rn(total);
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    if (c < C && j < M) { wgt[c] *= inv_total; lp += wgt[c]; }
  }
  float isum = lp;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, isum, o);
    if (lane >= o) isum += v;
  }
  float run = isum - lp;
  if (lane == 0) s_cdf[w][0] = 0.0f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    if (c < C && j < M) { run += wgt[c]; s_cdf[w][j + 1] = run; }
  }
  __syncwarp();
  if (cdf_out) for (int j = lane; j < S; j += 32) cdf_out[n * S + j] = s_cdf[w][j];
  for (int k = lane; k < K; k += 32) {
    int64_t ind;
    new_z[n * K + k] = invert_cdf(s_cdf[w], Z, S, K, k, &ind);
    if (inds_out) inds_out[n * K + k] = ind;
  }
}

// ------------------------------------------------------------------------------------ merge (cat_z_vals)
// One warp per ray.  Every global load of the ray (old depths and sdf, new depths and sdf) is issued before anything is used:
// with the loads inside the rank loops a warp paid ~10 dependent HBM latencies per ray and the kernel ran at 37 % of the copy
// bandwidth with every SM full of waiting warps.  MAXC = old samples per lane (4: S <= 128, 8: S <= kMaxS); K <= 64.
template <int MAXC>
__global__ void merge_z_kernel(const float* __restrict__ z, const float* __restrict__ nz, const float* __restrict__ sdf,
                               const float* __restrict__ nsdf, int64_t N, int S, int K, float* __restrict__ z_out,
                               float* __restrict__ sdf_out) {
  __shared__ float s_z[kWarpsPerBlock][kMaxS], s_n[kWarpsPerBlock][64];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + w;
  if (n >= N) return;
  float zv[MAXC], sv[MAXC], nv[2], nsv[2];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = lane + 32 * c;
    zv[c] = j < S ? z[n * S + j] : 0.0f;
    sv[c] = (sdf_out && j < S) ? sdf[n * S + j] : 0.0f;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int k = lane + 32 * c;
    nv[c] = k < K ? nz[n * K + k] : 0.0f;
    nsv[c] = (sdf_out && k < K) ? nsdf[n * K + k] : 0.0f;
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c)
    if (lane + 32 * c < S) s_z[w][lane + 32 * c] = zv[c];
#pragma unroll
  for (int c = 0; c < 2; ++c)
    if (lane + 32 * c < K) s_n[w][lane + 32 * c] = nv[c];
  __syncwarp();
  const int T = S + K;
  // The new depths of up_sample are inverse-CDF samples of increasing u: already sorted.  Then "# new < v" is a lower bound
  // (log2 K steps instead of K compares per old element; the kernel is issue-bound, ncu: 96 % of the issue slots busy).
  // Any other input (the generic C-ABI contract) takes the counting loop.
  bool sorted_new = true;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int k = lane + 32 * c;
    if (k + 1 < K) sorted_new = sorted_new && (s_n[w][k] <= s_n[w][k + 1]);
  }
  sorted_new = __all_sync(0xffffffffu, sorted_new);
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {             // old element: ties go before new ones
    const int j = lane + 32 * c;
    if (j < S) {
      const float v = zv[c];
      int r = j;
      if (sorted_new) {
        int lo = 0, hi = K;                    // first k with s_n[k] >= v  ==  # new < v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_n[w][mid] < v) lo = mid + 1; else hi = mid; }
        r += lo;
      } else {
        for (int k = 0; k < K; ++k) r += s_n[w][k] < v;
      }
      z_out[n * T + r] = v;
      if (sdf_out) sdf_out[n * T + r] = sv[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int k = lane + 32 * c;
    if (k < K) {
      const float v = nv[c];
      int lo = 0, hi = S;                      // # old <= v
      while (lo < hi) { int mid = (lo + hi) >> 1; if (s_z[w][mid] <= v) lo = mid + 1; else hi = mid; }
      int r = lo;
      if (sorted_new) {
        // rank among the (sorted) new depths with ties broken by index: every q < k is <= v, every q > k is >= v, so only
        // equal neighbours after k could differ from k itself -- and they are not counted (u == v needs q < k)
        int e = k;                             // # of q with (u < v) or (u == v and q < k) == k - (# q < k with u > v) == k
        r += e;
      } else {
        for (int q = 0; q < K; ++q) { float u = s_n[w][q]; r += (u < v) || (u == v && q < k); }
      }
      z_out[n * T + r] = v;
      if (sdf_out) sdf_out[n * T + r] = nsv[c];
    }
  }
}

// ------------------------------------------------------------------------------------ compositing
// neus_renderer.py:360-420.  C = samples per lane (S = 32*C exactly, or ragged via the generic path).
struct CompositeIn {
  const float* sdf; const float4* grad; const float* rgb; const float* z; const float* dists;
  const float* rays_d; const float* rays_d_norm; const float* variance;
  float cos_anneal; int eval_mode; int64_t N; int S;
};

__device__ __forceinline__ float inv_s_of(const float* variance, bool* clipped) {
  float s = expf(variance[0] * 10.0f);
  *clipped = s < 1e-3f || s > 1e3f;
  return fminf(fmaxf(s, 1e-3f), 1e3f);
}

template <int MAXC>
__device__ __forceinline__ void composite_alpha(const CompositeIn& a, int64_t base, int lane, int C, float3 dir,
                                                float inv_s, float (&alpha)[MAXC], float (&tc)[MAXC],
                                                float (&pc)[MAXC], float (&nc)[MAXC], float& prod) {
  prod = 1.0f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    alpha[c] = 0.0f; tc[c] = 0.0f; pc[c] = 0.0f; nc[c] = 0.0f;
    if (c < C && j < a.S) {
      float4 g = a.grad[base + j];
      float f = a.sdf[base + j], dist = a.dists[base + j];
      tc[c] = dir.x * g.x + dir.y * g.y + dir.z * g.z;
      float ic = -(fmaxf(-tc[c] * 0.5f + 0.5f, 0.0f) * (1.0f - a.cos_anneal) + fmaxf(-tc[c], 0.0f) * a.cos_anneal);
      float half = ic * dist * 0.5f;
      pc[c] = sigmoidf_((f - half) * inv_s);
      nc[c] = sigmoidf_((f + half) * inv_s);
      float ar = (pc[c] - nc[c] + 1e-5f) / (pc[c] + 1e-5f);
      alpha[c] = fminf(fmaxf(ar, 0.0f), 1.0f);
      prod *= (1.0f - alpha[c] + 1e-7f);
    }
  }
}

__device__ __forceinline__ float warp_excl_prod(float prod, int lane) {
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= v;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 1.0f : T;
}

template <int MAXC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const CompositeIn a, float* __restrict__ weights, float* __restrict__ color,
                     float* __restrict__ depth, float* __restrict__ wz, float* __restrict__ cdf, float* __restrict__ wsum,
                     float* __restrict__ wmax, float* __restrict__ inv_s_out) {
  int lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int C = (a.S + 31) / 32;
  const int64_t base = n * a.S;
  bool clipped;
  float inv_s = inv_s_of(a.variance, &clipped);
  if (n == 0 && lane == 0 && inv_s_out) inv_s_out[0] = inv_s;
  float3 dir = make_float3(a.rays_d[n * 3], a.rays_d[n * 3 + 1], a.rays_d[n * 3 + 2]);
  float alpha[MAXC], tc[MAXC], pc[MAXC], nc[MAXC], prod;
  composite_alpha<MAXC>(a, base, lane, C, dir, inv_s, alpha, tc, pc, nc, prod);
  float T = warp_excl_prod(prod, lane);
  float cr = 0, cg = 0, cb = 0, dp = 0, ws = 0, wm = 0;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    if (c < C && j < a.S) {
      float wv = alpha[c] * T;
      T *= (1.0f - alpha[c] + 1e-7f);
      weights[base + j] = wv;
      if (cdf) cdf[base + j] = pc[c];
      const float* col = a.rgb + (base + j) * 3;
      cr += wv * col[0]; cg += wv * col[1]; cb += wv * col[2];
      dp += wv * a.z[base + j];
      ws += wv; wm = fmaxf(wm, wv);
    }
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dp = warp_sum(dp); ws = warp_sum(ws); wm = warp_max(wm);
  if (lane == 0) {
    color[n * 3] = cr; color[n * 3 + 1] = cg; color[n * 3 + 2] = cb;
    depth[n] = a.eval_mode ? dp / a.rays_d_norm[n] : dp;
    if (wz) wz[n] = dp;
    if (wsum) wsum[n] = ws;
    if (wmax) wmax[n] = wm;
  }
}

template <int MAXC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const CompositeIn a, const float* __restrict__ d_color, const float* __restrict__ d_depth,
                     const float* __restrict__ d_weights, const float4* d_grad_in, float* __restrict__ d_sdf,
                     float4* d_grad, float* __restrict__ d_rgb, float* __restrict__ d_variance,
                     float* __restrict__ d_rays_d) {
  int lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int C = (a.S + 31) / 32;
  const int64_t base = n * a.S;
  bool clipped;
  float inv_s = inv_s_of(a.variance, &clipped);
  float3 dir = make_float3(a.rays_d[n * 3], a.rays_d[n * 3 + 1], a.rays_d[n * 3 + 2]);
  float alpha[MAXC], tc[MAXC], pc[MAXC], nc[MAXC], prod;
  composite_alpha<MAXC>(a, base, lane, C, dir, inv_s, alpha, tc, pc, nc, prod);
  float T0 = warp_excl_prod(prod, lane);
  float dcr = d_color ? d_color[n * 3] : 0.0f, dcg = d_color ? d_color[n * 3 + 1] : 0.0f,
        dcb = d_color ? d_color[n * 3 + 2] : 0.0f;
  float ddp = d_depth ? d_depth[n] : 0.0f;
  if (a.eval_mode) ddp /= a.rays_d_norm[n];
  // pass 1: weights, dL/dw, per-lane sum of dw*w
  float Tc[MAXC], dw[MAXC], wv[MAXC];
  float T = T0, lsum = 0.0f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    int j = lane * C + c;
    Tc[c] = 0; dw[c] = 0; wv[c] = 0;
    if (c < C && j < a.S) {
      Tc[c] = T;
      wv[c] = alpha[c] * T;
      T *= (1.0f - alpha[c] + 1e-7f);
      const float* col = a.rgb + (base + j) * 3;
      dw[c] = (d_weights ? d_weights[base + j] : 0.0f) + dcr * col[0] + dcg * col[1] + dcb * col[2] + ddp * a.z[base + j];
      float* o = d_rgb + (base + j) * 3;
      o[0] = wv[c] * dcr; o[1] = wv[c] * dcg; o[2] = wv[c] * dcb;
      lsum += dw[c] * wv[c];
    }
  }
  // exclusive suffix sum over lanes
  float incl = lsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float v = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += v;
  }
  float R = incl - lsum;     // sum over lanes > this one
  float dvar = 0.0f, ddx = 0.0f, ddy = 0.0f, ddz = 0.0f;
#pragma unroll
  for (int c = MAXC - 1; c >= 0; --c) {
    int j = lane * C + c;
    if (c < C && j < a.S) {
      float one_m = 1.0f - alpha[c] + 1e-7f;
      float dalpha = dw[c] * Tc[c] - R / one_m;
      R += dw[c] * wv[c];
      float f = a.sdf[base + j], dist = a.dists[base + j];
      float4 g = a.grad[base + j];
      float den = pc[c] + 1e-5f;
      float ar = (pc[c] - nc[c] + 1e-5f) / den;
      float dar = (ar >= 0.0f && ar <= 1.0f) ? dalpha : 0.0f;
      float dpc = dar * (nc[c] / (den * den));
      float dnc = -dar / den;
      float gp = dpc * pc[c] * (1.0f - pc[c]);      // d/d(ep*s)
      float gn = dnc * nc[c] * (1.0f - nc[c]);      // d/d(en*s)
      float r = a.cos_anneal;
      float t = tc[c];
      float ic = -(fmaxf(-t * 0.5f + 0.5f, 0.0f) * (1.0f - r) + fmaxf(-t, 0.0f) * r);
      float half = ic * dist * 0.5f;
      float ep = f - half, en = f + half;
      dvar += gp * ep + gn * en;
      float dep = gp * inv_s, den_ = gn * inv_s;
      d_sdf[base + j] = dep + den_;
      float dic = (den_ - dep) * dist * 0.5f;
      float dtc = dic * (((-t * 0.5f + 0.5f) > 0.0f ? 0.5f * (1.0f - r) : 0.0f) + ((-t) > 0.0f ? r : 0.0f));
      // write-only unless the caller has upstream gradients of normals / sdf_flows (d_grad_in may alias d_grad: each
      // element is read and written by the same thread)
      float4 og = d_grad_in ? d_grad_in[base + j] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      og.x += dtc * dir.x; og.y += dtc * dir.y; og.z += dtc * dir.z;
      d_grad[base + j] = og;
      ddx += dtc * g.x; ddy += dtc * g.y; ddz += dtc * g.z;
    }
  }
  dvar = warp_sum(dvar); ddx = warp_sum(ddx); ddy = warp_sum(ddy); ddz = warp_sum(ddz);
  if (lane == 0) {
    if (d_rays_d) { d_rays_d[n * 3] = ddx; d_rays_d[n * 3 + 1] = ddy; d_rays_d[n * 3 + 2] = ddz; }
    if (d_variance && !clipped) atomicAdd(d_variance, dvar * 10.0f * inv_s);
  }
}


// ---- S == 128 fast path (the shipped 64 + 64 samples): lane l owns samples 4l .. 4l+3, so every per-sample array is read
// and written with 16-byte vector accesses that are contiguous across the warp (one 512-byte request per scalar array instead
// of four 25 %-efficient 4-byte gathers; ncu counted 74 % excessive sectors in the scalar kernel), every load of the ray is in
// flight before the first use, and the backward keeps what it loaded instead of reading it again for its second pass.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void to_arr(const float4 v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
struct Alpha4 { float alpha[4], tc[4], pc[4], nc[4], prod; };
__device__ __forceinline__ Alpha4 alpha4(const float (&f)[4], const float (&dist)[4], const float4 (&g)[4], float3 dir, float inv_s,
                                          float cos_anneal) {
  Alpha4 A;
  A.prod = 1.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    A.tc[c] = dir.x * g[c].x + dir.y * g[c].y + dir.z * g[c].z;
    const float ic = -(fmaxf(-A.tc[c] * 0.5f + 0.5f, 0.0f) * (1.0f - cos_anneal) + fmaxf(-A.tc[c], 0.0f) * cos_anneal);
    const float half = ic * dist[c] * 0.5f;
    A.pc[c] = sigmoidf_((f[c] - half) * inv_s);
    A.nc[c] = sigmoidf_((f[c] + half) * inv_s);
    const float ar = (A.pc[c] - A.nc[c] + 1e-5f) / (A.pc[c] + 1e-5f);
    A.alpha[c] = fminf(fmaxf(ar, 0.0f), 1.0f);
    A.prod *= (1.0f - A.alpha[c] + 1e-7f);
  }
  return A;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd128_kernel(const CompositeIn a, float* __restrict__ weights, float* __restrict__ color, float* __restrict__ depth,
                        float* __restrict__ wz, float* __restrict__ cdf, float* __restrict__ wsum, float* __restrict__ wmax,
                        float* __restrict__ inv_s_out) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t b4 = n * 128 + lane * 4;
  float f[4], dist[4], zz[4], col[12];
  float4 g[4];
  to_arr(ldg4(a.sdf + b4), f); to_arr(ldg4(a.dists + b4), dist); to_arr(ldg4(a.z + b4), zz);
#pragma unroll
  for (int c = 0; c < 4; ++c) g[c] = __ldg(a.grad + b4 + c);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 v = ldg4(a.rgb + b4 * 3 + 4 * k);
    col[4 * k] = v.x; col[4 * k + 1] = v.y; col[4 * k + 2] = v.z; col[4 * k + 3] = v.w;
  }
  bool clipped;
  const float inv_s = inv_s_of(a.variance, &clipped);
  if (n == 0 && lane == 0 && inv_s_out) inv_s_out[0] = inv_s;
  const float3 dir = make_float3(a.rays_d[n * 3], a.rays_d[n * 3 + 1], a.rays_d[n * 3 + 2]);
  const Alpha4 A = alpha4(f, dist, g, dir, inv_s, a.cos_anneal);
  float T = warp_excl_prod(A.prod, lane);
  float cr = 0, cg = 0, cb = 0, dp = 0, ws = 0, wm = 0, wv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    wv[c] = A.alpha[c] * T;
    T *= (1.0f - A.alpha[c] + 1e-7f);
    cr += wv[c] * col[3 * c]; cg += wv[c] * col[3 * c + 1]; cb += wv[c] * col[3 * c + 2];
    dp += wv[c] * zz[c];
    ws += wv[c]; wm = fmaxf(wm, wv[c]);
  }
  *reinterpret_cast<float4*>(weights + b4) = make_float4(wv[0], wv[1], wv[2], wv[3]);
  if (cdf) *reinterpret_cast<float4*>(cdf + b4) = make_float4(A.pc[0], A.pc[1], A.pc[2], A.pc[3]);
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dp = warp_sum(dp); ws = warp_sum(ws); wm = warp_max(wm);
  if (lane == 0) {
    color[n * 3] = cr; color[n * 3 + 1] = cg; color[n * 3 + 2] = cb;
    depth[n] = a.eval_mode ? dp / a.rays_d_norm[n] : dp;
    if (wz) wz[n] = dp;
    if (wsum) wsum[n] = ws;
    if (wmax) wmax[n] = wm;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd128_kernel(const CompositeIn a, const float* __restrict__ d_color, const float* __restrict__ d_depth,
                        const float* __restrict__ d_weights, const float4* d_grad_in, float* __restrict__ d_sdf, float4* d_grad,
                        float* __restrict__ d_rgb, float* __restrict__ d_variance, float* __restrict__ d_rays_d) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t b4 = n * 128 + lane * 4;
  float f[4], dist[4], zz[4], col[12], dwu[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  float4 g[4], gin[4];
  to_arr(ldg4(a.sdf + b4), f); to_arr(ldg4(a.dists + b4), dist); to_arr(ldg4(a.z + b4), zz);
  if (d_weights) to_arr(ldg4(d_weights + b4), dwu);
#pragma unroll
  for (int c = 0; c < 4; ++c) g[c] = __ldg(a.grad + b4 + c);
#pragma unroll
  for (int c = 0; c < 4; ++c) gin[c] = d_grad_in ? d_grad_in[b4 + c] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 v = ldg4(a.rgb + b4 * 3 + 4 * k);
    col[4 * k] = v.x; col[4 * k + 1] = v.y; col[4 * k + 2] = v.z; col[4 * k + 3] = v.w;
  }
  bool clipped;
  const float inv_s = inv_s_of(a.variance, &clipped);
  const float3 dir = make_float3(a.rays_d[n * 3], a.rays_d[n * 3 + 1], a.rays_d[n * 3 + 2]);
  const Alpha4 A = alpha4(f, dist, g, dir, inv_s, a.cos_anneal);
  const float T0 = warp_excl_prod(A.prod, lane);
  const float dcr = d_color ? d_color[n * 3] : 0.0f, dcg = d_color ? d_color[n * 3 + 1] : 0.0f, dcb = d_color ? d_color[n * 3 + 2] : 0.0f;
  float ddp = d_depth ? d_depth[n] : 0.0f;
  if (a.eval_mode) ddp /= a.rays_d_norm[n];
  // pass 1: weights, dL/dw, per-lane sum of dw * w, d_rgb
  float Tc[4], dw[4], wv[4], drgb[12];
  float T = T0, lsum = 0.0f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    Tc[c] = T;
    wv[c] = A.alpha[c] * T;
    T *= (1.0f - A.alpha[c] + 1e-7f);
    dw[c] = dwu[c] + dcr * col[3 * c] + dcg * col[3 * c + 1] + dcb * col[3 * c + 2] + ddp * zz[c];
    drgb[3 * c] = wv[c] * dcr; drgb[3 * c + 1] = wv[c] * dcg; drgb[3 * c + 2] = wv[c] * dcb;
    lsum += dw[c] * wv[c];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    *reinterpret_cast<float4*>(d_rgb + b4 * 3 + 4 * k) = make_float4(drgb[4 * k], drgb[4 * k + 1], drgb[4 * k + 2], drgb[4 * k + 3]);
  float incl = lsum;                               // exclusive suffix sum over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += v;
  }
  float R = incl - lsum;
  float dvar = 0.0f, ddx = 0.0f, ddy = 0.0f, ddz = 0.0f, dsdf[4];
#pragma unroll
  for (int c = 3; c >= 0; --c) {
    const float one_m = 1.0f - A.alpha[c] + 1e-7f;
    const float dalpha = dw[c] * Tc[c] - R / one_m;
    R += dw[c] * wv[c];
    const float den = A.pc[c] + 1e-5f;
    const float ar = (A.pc[c] - A.nc[c] + 1e-5f) / den;
    const float dar = (ar >= 0.0f && ar <= 1.0f) ? dalpha : 0.0f;
    const float dpc = dar * (A.nc[c] / (den * den));
    const float dnc = -dar / den;
    const float gp = dpc * A.pc[c] * (1.0f - A.pc[c]);
    const float gn = dnc * A.nc[c] * (1.0f - A.nc[c]);
    const float r = a.cos_anneal, t = A.tc[c];
    const float ic = -(fmaxf(-t * 0.5f + 0.5f, 0.0f) * (1.0f - r) + fmaxf(-t, 0.0f) * r);
    const float half = ic * dist[c] * 0.5f;
    dvar += gp * (f[c] - half) + gn * (f[c] + half);
    const float dep = gp * inv_s, den_ = gn * inv_s;
    dsdf[c] = dep + den_;
    const float dic = (den_ - dep) * dist[c] * 0.5f;
    const float dtc = dic * (((-t * 0.5f + 0.5f) > 0.0f ? 0.5f * (1.0f - r) : 0.0f) + ((-t) > 0.0f ? r : 0.0f));
    gin[c].x += dtc * dir.x; gin[c].y += dtc * dir.y; gin[c].z += dtc * dir.z;
    ddx += dtc * g[c].x; ddy += dtc * g[c].y; ddz += dtc * g[c].z;
  }
  *reinterpret_cast<float4*>(d_sdf + b4) = make_float4(dsdf[0], dsdf[1], dsdf[2], dsdf[3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) d_grad[b4 + c] = gin[c];
  dvar = warp_sum(dvar); ddx = warp_sum(ddx); ddy = warp_sum(ddy); ddz = warp_sum(ddz);
  if (lane == 0) {
    if (d_rays_d) { d_rays_d[n * 3] = ddx; d_rays_d[n * 3 + 1] = ddy; d_rays_d[n * 3 + 2] = ddz; }
    if (d_variance && !clipped) atomicAdd(d_variance, dvar * 10.0f * inv_s);
  }
}

}  // namespace cope

using namespace cope;

static inline dim3 grid1d(int64_t n, int bs = 256) { return dim3((unsigned)ceil_div(n, bs)); }
static inline dim3 grid_rays(int64_t N) { return dim3((unsigned)ceil_div(N, kWarpsPerBlock)); }

extern "C" {

int cope_coarse_z(const float* near, const float* far, const float* t_rand, int64_t N, int S, float* z, cope_stream_t s) {
  if (N <= 0) return 0;
  coarse_z_kernel<<<grid1d(N * S), 256, 0, as_stream(s)>>>(near, far, t_rand, N, S, z);
  COPE_CHECK_LAUNCH("coarse_z");
  return 0;
}

int cope_ray_points(const float* rays_o, const float* rays_d, const float* z, const float* time_step, const float* near,
                    const float* far, int n_coarse, int64_t N, int S, int use_mid, float* pts_time, float* dists,
                    float* mid_z, cope_stream_t s) {
  if (N <= 0) return 0;
  ray_points_kernel<<<grid1d(N * S), 256, 0, as_stream(s)>>>(rays_o, rays_d, z, time_step, near, far, n_coarse, N, S,
                                                            use_mid, reinterpret_cast<float4*>(pts_time), dists, mid_z);
  COPE_CHECK_LAUNCH("ray_points");
  return 0;
}

int cope_ray_points_bwd(const float* d_pts, const float* mid_z, const float* d_dirs_pp, int64_t N, int S,
                        float* d_rays_o, float* d_rays_d, cope_stream_t s) {
  if (N <= 0) return 0;
  ray_points_bwd_kernel<<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(
      reinterpret_cast<const float4*>(d_pts), mid_z, d_dirs_pp, N, S, d_rays_o, d_rays_d);
  COPE_CHECK_LAUNCH("ray_points_bwd");
  return 0;
}

int cope_sample_cdf(const float* cdf, const float* bins, int64_t N, int S, int K, float* samples, int64_t* inds,
                    cope_stream_t s) {
  COPE_REQUIRE(S >= 1 && S <= kMaxS, "sample_cdf: S=%d outside [1,%d]", S, kMaxS);
  if (N <= 0) return 0;
  sample_cdf_kernel<<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(cdf, bins, N, S, K, samples, inds);
  COPE_CHECK_LAUNCH("sample_cdf");
  return 0;
}

int cope_upsample(const float* z, const float* sdf, int64_t N, int S, int K, float inv_s, float* new_z, float* cdf_out,
                  int64_t* inds_out, cope_stream_t s) {
  COPE_REQUIRE(S >= 2 && S <= kMaxS, "upsample: S=%d outside [2,%d]", S, kMaxS);
  if (N <= 0) return 0;
  if (S - 1 <= 128)
    upsample_kernel<4><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(z, sdf, N, S, K, inv_s, new_z, cdf_out, inds_out);
  else
    upsample_kernel<8><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(z, sdf, N, S, K, inv_s, new_z, cdf_out, inds_out);
  COPE_CHECK_LAUNCH("upsample");
  return 0;
}

int cope_merge_z(const float* z, const float* new_z, const float* sdf, const float* new_sdf, int64_t N, int S, int K,
                 float* z_out, float* sdf_out, cope_stream_t s) {
  COPE_REQUIRE(S >= 1 && S <= kMaxS && K >= 1 && K <= 64, "merge_z: S=%d K=%d out of range", S, K);
  if (N <= 0) return 0;
  if (S <= 128)
    merge_z_kernel<4><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(z, new_z, sdf, new_sdf, N, S, K, z_out,
                                                                              sdf ? sdf_out : nullptr);
  else
    merge_z_kernel<8><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(z, new_z, sdf, new_sdf, N, S, K, z_out,
                                                                              sdf ? sdf_out : nullptr);
  COPE_CHECK_LAUNCH("merge_z");
  return 0;
}

int cope_composite_fwd(const float* sdf, const float* grad, const float* rgb, const float* z, const float* dists,
                       const float* rays_d, const float* rays_d_norm, const float* variance, float cos_anneal,
                       int eval_mode, int64_t N, int S, float* weights, float* color, float* depth, float* weighted_z,
                       float* cdf, float* wsum, float* wmax, float* inv_s_out, cope_stream_t s) {
  COPE_REQUIRE(S >= 1 && S <= 256, "composite: S=%d outside [1,256]", S);
  if (N <= 0) return 0;
  CompositeIn a{sdf, reinterpret_cast<const float4*>(grad), rgb, z, dists, rays_d, rays_d_norm, variance,
                cos_anneal, eval_mode, N, S};
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(z) |
                           reinterpret_cast<uintptr_t>(dists) | reinterpret_cast<uintptr_t>(weights) | reinterpret_cast<uintptr_t>(cdf)) & 15) == 0;
  if (S == 128 && aligned16 && !getenv("COPE_COMPOSITE_SCALAR"))
    composite_fwd128_kernel<<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(a, weights, color, depth, weighted_z, cdf, wsum,
                                                                                   wmax, inv_s_out);
  else if (S <= 128)
    composite_fwd_kernel<4><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(a, weights, color, depth, weighted_z, cdf,
                                                                                   wsum, wmax, inv_s_out);
  else
    composite_fwd_kernel<8><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(a, weights, color, depth, weighted_z, cdf,
                                                                                   wsum, wmax, inv_s_out);
  COPE_CHECK_LAUNCH("composite_fwd");
  return 0;
}

int cope_composite_bwd(const float* sdf, const float* grad, const float* rgb, const float* z, const float* dists,
                       const float* rays_d, const float* rays_d_norm, const float* variance, float cos_anneal,
                       int eval_mode, int64_t N, int S, const float* d_color, const float* d_depth,
                       const float* d_weights, const float* d_grad_in, float* d_sdf, float* d_grad, float* d_rgb,
                       float* d_variance, float* d_rays_d, cope_stream_t s) {
  COPE_REQUIRE(S >= 1 && S <= 256, "composite: S=%d outside [1,256]", S);
  if (N <= 0) return 0;
  CompositeIn a{sdf, reinterpret_cast<const float4*>(grad), rgb, z, dists, rays_d, rays_d_norm, variance,
                cos_anneal, eval_mode, N, S};
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(sdf) | reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(z) |
                           reinterpret_cast<uintptr_t>(dists) | reinterpret_cast<uintptr_t>(d_weights) | reinterpret_cast<uintptr_t>(d_sdf) |
                           reinterpret_cast<uintptr_t>(d_rgb)) & 15) == 0;
  if (S == 128 && aligned16 && !getenv("COPE_COMPOSITE_SCALAR"))
    composite_bwd128_kernel<<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(
        a, d_color, d_depth, d_weights, reinterpret_cast<const float4*>(d_grad_in), d_sdf, reinterpret_cast<float4*>(d_grad), d_rgb,
        d_variance, d_rays_d);
  else if (S <= 128)
    composite_bwd_kernel<4><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(
        a, d_color, d_depth, d_weights, reinterpret_cast<const float4*>(d_grad_in), d_sdf, reinterpret_cast<float4*>(d_grad),
        d_rgb, d_variance, d_rays_d);
  else
    composite_bwd_kernel<8><<<grid_rays(N), kWarpsPerBlock * 32, 0, as_stream(s)>>>(
        a, d_color, d_depth, d_weights, reinterpret_cast<const float4*>(d_grad_in), d_sdf, reinterpret_cast<float4*>(d_grad),
        d_rgb, d_variance, d_rays_d);
  COPE_CHECK_LAUNCH("composite_bwd");
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------ evaluation-render reductions
// model/training.py:236-283: one warp per ray; weighted normal sum + arg-max-weight sample depth in the camera frame + (optional)
// the predicted forward optical flow.  The reference integrates the scene flow of EVERY sample point over the sub-steps
// (p <- p + dt (w_t x p + v_t), :269-272) and then takes the weight average (:273-275); each sub-step is affine in p, so the
// whole integration is one 3 x 4 affine map F and  sum_s w (F [p; 1]) = F [sum_s w p; sum_s w]  - four numbers per ray.
namespace cope {
struct FlowMapArgs {
  const float* F;          // [12] row-major 3 x 4 affine scene-flow map (null: no flow output)
  const float* KS;         // [9]  scale_mat[:3,:3] @ camera_mat[:3,:3]
  const float* pix;        // [N x 2] normalised pixel of each ray
  float sx, sy;            // w / 2, h / 2 (:296-297)
  float* flow;             // [N x 2] flow in pixels
};
__global__ void eval_reduce_kernel(const float* __restrict__ w, const float4* __restrict__ grad, const float4* __restrict__ pts,
                                   const float* __restrict__ M, int64_t N, int S, float* __restrict__ nrm, float* __restrict__ dhw,
                                   const FlowMapArgs fa) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= N) return;
  float ax = 0.0f, ay = 0.0f, az = 0.0f, best = -1.0f;
  float px = 0.0f, py = 0.0f, pz = 0.0f, pw = 0.0f;
  int besti = 0;
  for (int s = lane; s < S; s += 32) {
    const float wv = w[n * S + s];
    const float4 g = grad[n * S + s];
    ax = fmaf(wv, g.x, ax); ay = fmaf(wv, g.y, ay); az = fmaf(wv, g.z, az);
    if (fa.F) {
      const float4 p = pts[n * S + s];
      px = fmaf(wv, p.x, px); py = fmaf(wv, p.y, py); pz = fmaf(wv, p.z, pz); pw += wv;
    }
    if (wv > best) { best = wv; besti = s; }          // first maximum within the lane (torch.max returns the first index)
  }
  ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
  if (fa.F) { px = warp_sum(px); py = warp_sum(py); pz = warp_sum(pz); pw = warp_sum(pw); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if (lane == 0) {
    nrm[n * 3 + 0] = M[0] * ax + M[1] * ay + M[2] * az;
    nrm[n * 3 + 1] = M[4] * ax + M[5] * ay + M[6] * az;
    nrm[n * 3 + 2] = M[8] * ax + M[9] * ay + M[10] * az;
    const float4 p = pts[n * S + besti];
    dhw[n] = -(M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11]);
    if (fa.F) {
      float m[3], q[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) m[r] = fa.F[r * 4] * px + fa.F[r * 4 + 1] * py + fa.F[r * 4 + 2] * pz + fa.F[r * 4 + 3] * pw;
#pragma unroll
      for (int r = 0; r < 3; ++r) q[r] = fa.KS[r * 3] * m[0] + fa.KS[r * 3 + 1] * m[1] + fa.KS[r * 3 + 2] * m[2];
      fa.flow[n * 2] = (q[0] / q[2] - fa.pix[n * 2]) * fa.sx;
      fa.flow[n * 2 + 1] = (q[1] / q[2] - fa.pix[n * 2 + 1]) * fa.sy;
    }
  }
}
}  // namespace cope

extern "C" int cope_eval_reduce(const float* weights, const float* grad, const float* pts, const float* world_mat, int64_t N, int S,
                                float* normal_out, float* depth_hw_out, const float* flow_affine, const float* KS,
                                const float* pix_norm, float flow_sx, float flow_sy, float* flow_out, cope_stream_t s) {
  using namespace cope;
  if (N <= 0) return 0;
  COPE_REQUIRE(S > 0 && weights && grad && pts && world_mat && normal_out && depth_hw_out, "eval_reduce: null argument");
  COPE_REQUIRE(!flow_affine || (KS && pix_norm && flow_out), "eval_reduce: the flow output needs KS, pix_norm and flow_out");
  FlowMapArgs fa{flow_affine, KS, pix_norm, flow_sx, flow_sy, flow_out};
  eval_reduce_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, as_stream(s)>>>(weights, reinterpret_cast<const float4*>(grad),
                                                                        reinterpret_cast<const float4*>(pts), world_mat, N, S,
                                                                        normal_out, depth_hw_out, fa);
  COPE_CHECK_LAUNCH("eval_reduce");
  return 0;
}
