// Loss reductions of the training step, each as ONE forward and ONE backward launch
// instead of the reference's chains of elementwise torch kernels:
//
//   step losses   rgb L1 (model/training.py:508) + eikonal (train.py:526) + SDF-flow loss (train.py:467-477)
//   weighted pts  sum_s w[n,s] * [p[n,s], 1]  -- the only per-sample part of the flow-RGB loss (train.py:488-489):
//                 sum_s w (R p + T) = R (sum_s w p) + T (sum_s w), so every reference frame reuses the same 4 numbers per ray
//   flow-RGB      projection into the reference frames, pixel flow, bilinear border-clamped warp (train.py:235-244 =
//                 grid_sample(align_corners=True)), masked L1 against the target colours (train.py:486-517)
//
// All are HBM-trivial (<= 48 B per sample, read once); the point is launch count: ~60 launches -> 2 per loss group.
#include <algorithm>

#include "common.cuh"

namespace cope {
namespace {

constexpr int kLossThreads = 256;

struct StepLossArgs {
  const float* color; const float* rgb_gt;          // [N x 3]
  const float4* grad4;                              // [P x 4] (normal | sdf flow)
  const float4* pts4;                               // [P x 4] (x, y, z, t)            (SDF-flow term)
  const float* weights;                             // [P]                              (SDF-flow term)
  const float* motion;                              // [6] angular velocity | velocity  (null: no SDF-flow term)
  const float* w_sum_global;                        // [1] or null: normaliser of the SDF-flow term summed over all ranks
  int64_t N, P;
  float w_rgb, w_eik, w_flow;
};

__device__ __forceinline__ float signf_(float x) { return (x > 0.0f) - (x < 0.0f); }

// block-wide sum of K per-thread values, result valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* sh /* [K * 8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * 8 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float x = lane < (int)(blockDim.x >> 5) ? sh[k * 8 + lane] : 0.0f;
      v[k] = warp_sum(x);
    }
  }
  __syncthreads();
}

// ws: [0] sum |rgb - gt|, [1] sum (|n| - 1)^2, [2] sum |flow.n + sdf_flow| w, [3] sum w, [4] block counter (uint)
__global__ void __launch_bounds__(kLossThreads) step_losses_fwd_kernel(const StepLossArgs a, float* __restrict__ ws,
                                                                      float* __restrict__ losses, float* __restrict__ coef) {
  __shared__ float sh[4 * 8];
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < a.N * 3; i += stride) acc[0] += fabsf(a.color[i] - a.rgb_gt[i]);
  float om[6] = {0, 0, 0, 0, 0, 0};
  if (a.motion)
#pragma unroll
    for (int k = 0; k < 6; ++k) om[k] = a.motion[k];
  for (int64_t p = i0; p < a.P; p += stride) {
    const float4 g = a.grad4[p];
    const float nn = sqrtf(g.x * g.x + g.y * g.y + g.z * g.z);
    acc[1] += (nn - 1.0f) * (nn - 1.0f);
    if (a.motion) {
      const float4 x = a.pts4[p];
      const float w = a.weights[p];
      const float sx = om[1] * x.z - om[2] * x.y + om[3], sy = om[2] * x.x - om[0] * x.z + om[4],
                  sz = om[0] * x.y - om[1] * x.x + om[5];
      acc[2] += fabsf(sx * g.x + sy * g.y + sz * g.z + g.w) * w;
      acc[3] += w;
    }
  }
  block_sum<4>(acc, sh);
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (acc[k] != 0.0f) atomicAdd(ws + k, acc[k]);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(ws + 4), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* v = ws;
    const float s_rgb = v[0], s_eik = v[1], s_flow = v[2];
    const float s_w = a.w_sum_global ? *a.w_sum_global : v[3];
    const float c_rgb = a.N > 0 ? 1.0f / (float)a.N : 0.0f, c_eik = a.P > 0 ? 1.0f / (float)a.P : 0.0f;
    const float c_flow = a.motion ? 1.0f / (s_w + 1e-10f) : 0.0f;
    const float l_rgb = s_rgb * c_rgb, l_eik = s_eik * c_eik, l_flow = s_flow * c_flow;
    losses[0] = a.w_rgb * l_rgb + a.w_eik * l_eik + a.w_flow * l_flow;
    losses[1] = l_rgb; losses[2] = l_eik; losses[3] = l_flow;
    coef[0] = a.w_rgb * c_rgb; coef[1] = a.w_eik * c_eik; coef[2] = a.w_flow * c_flow; coef[3] = s_w;
  }
}

// gradients of the step losses; coef = the three normalised loss weights written by the forward, g = dL/d total (device)
__global__ void __launch_bounds__(kLossThreads) step_losses_bwd_kernel(const StepLossArgs a, const float* __restrict__ coef,
                                                                      const float* __restrict__ g_ptr, float* __restrict__ d_color,
                                                                      float4* __restrict__ d_grad4, float4* __restrict__ d_pts4,
                                                                      float* __restrict__ d_motion) {
  __shared__ float sh[6 * 8];
  const float g = g_ptr ? *g_ptr : 1.0f;
  const float c_rgb = coef[0] * g, c_eik = coef[1] * g, c_flow = coef[2] * g;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d_color)
    for (int64_t i = i0; i < a.N * 3; i += stride) d_color[i] = c_rgb * signf_(a.color[i] - a.rgb_gt[i]);
  float om[6] = {0, 0, 0, 0, 0, 0};
  if (a.motion)
#pragma unroll
    for (int k = 0; k < 6; ++k) om[k] = a.motion[k];
  float dm[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t p = i0; p < a.P; p += stride) {
    const float4 gr = a.grad4[p];
    const float nn = sqrtf(gr.x * gr.x + gr.y * gr.y + gr.z * gr.z);
    const float e = nn > 0.0f ? c_eik * 2.0f * (nn - 1.0f) / nn : 0.0f;      // torch's norm backward: 0 at the origin
    float4 dg = make_float4(e * gr.x, e * gr.y, e * gr.z, 0.0f);
    float4 dp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (a.motion) {
      const float4 x = a.pts4[p];
      const float sx = om[1] * x.z - om[2] * x.y + om[3], sy = om[2] * x.x - om[0] * x.z + om[4],
                  sz = om[0] * x.y - om[1] * x.x + om[5];
      const float s = signf_(sx * gr.x + sy * gr.y + sz * gr.z + gr.w) * a.weights[p] * c_flow;
      dg.x += s * sx; dg.y += s * sy; dg.z += s * sz; dg.w = s;
      const float ax = s * gr.x, ay = s * gr.y, az = s * gr.z;                // d / d scene_flow
      dp.x = ay * om[2] - az * om[1]; dp.y = az * om[0] - ax * om[2]; dp.z = ax * om[1] - ay * om[0];   // a x omega
      dm[0] += x.y * az - x.z * ay; dm[1] += x.z * ax - x.x * az; dm[2] += x.x * ay - x.y * ax;          // p x a
      dm[3] += ax; dm[4] += ay; dm[5] += az;
    }
    d_grad4[p] = dg;
    if (d_pts4) d_pts4[p] = dp;
  }
  if (a.motion && d_motion) {
    block_sum<6>(dm, sh);
    if (threadIdx.x == 0)
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (dm[k] != 0.0f) atomicAdd(d_motion + k, dm[k]);
  }
}

// ------------------------------------------------------------------------------------------------ weighted points
// one warp per ray: wp[n] = (sum_s w p_x, sum_s w p_y, sum_s w p_z, sum_s w)
__global__ void __launch_bounds__(256) weighted_points_fwd_kernel(const float* __restrict__ w, const float4* __restrict__ pts,
                                                                  int64_t N, int S, float4* __restrict__ wp) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float ax = 0, ay = 0, az = 0, aw = 0;
  for (int j = lane; j < S; j += 32) {
    const float ww = w[n * S + j];
    const float4 p = pts[n * S + j];
    ax += ww * p.x; ay += ww * p.y; az += ww * p.z; aw += ww;
  }
  ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az); aw = warp_sum(aw);
  if (lane == 0) wp[n] = make_float4(ax, ay, az, aw);
}

__global__ void __launch_bounds__(256) weighted_points_bwd_kernel(const float* __restrict__ w, const float4* __restrict__ pts,
                                                                  const float4* __restrict__ d_wp, int64_t P, int S,
                                                                  float* __restrict__ d_w, float4* __restrict__ d_pts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 d = d_wp[i / S];
  const float4 p = pts[i];
  if (d_w) d_w[i] = d.x * p.x + d.y * p.y + d.z * p.z + d.w;
  if (d_pts) {
    const float ww = w[i];
    d_pts[i] = make_float4(ww * d.x, ww * d.y, ww * d.z, 0.0f);
  }
}

// ------------------------------------------------------------------------------------------------ flow-RGB loss
struct FlowRgbArgs {
  const float4* wp;        // [N] weighted points (xyz, sum w)
  const float* w2c;        // [T x 16] row-major world -> reference-camera maps
  const float* KS;         // [T x 9]  scale_mat[:3,:3] @ ref_camera_mat[:3,:3]
  const float* npix;       // [N x 2] normalised pixel of each ray (x, y)
  const float* pix;        // [N x 2] pixel coordinates of each ray (x, y)
  const float* ref;        // [T x 3 x H x W] reference frames
  const float* rgb_gt;     // [N x 3]
  int64_t N; int T, H, W;
};

struct WarpGeom {            // per (ray, frame): projection + bilinear taps
  float m[3], q[3];
  float cx, cy;
  bool valid;
  float ix, iy, mx, my;      // clamped source coordinate and the clamp's derivative (0 / 1)
  int x0, y0;
};

// PyTorch grid_sample, padding_mode='border': clip_coordinates_set_grad
__device__ __forceinline__ float clip_border(float v, int size, float* mult) {
  if (v <= 0.0f) { *mult = 0.0f; return 0.0f; }
  const float mx = (float)(size - 1);
  if (v >= mx) { *mult = 0.0f; return mx; }
  *mult = 1.0f;
  return v;
}

__device__ __forceinline__ WarpGeom flow_geom(const FlowRgbArgs& a, int64_t n, int t) {
  WarpGeom G;
  const float4 w = a.wp[n];
  const float* M = a.w2c + t * 16;
  const float* K = a.KS + t * 9;
#pragma unroll
  for (int r = 0; r < 3; ++r) G.m[r] = M[r * 4] * w.x + M[r * 4 + 1] * w.y + M[r * 4 + 2] * w.z + M[r * 4 + 3] * w.w;
#pragma unroll
  for (int r = 0; r < 3; ++r) G.q[r] = K[r * 3] * G.m[0] + K[r * 3 + 1] * G.m[1] + K[r * 3 + 2] * G.m[2];
  const float u = G.q[0] / G.q[2], v = G.q[1] / G.q[2];
  const float fx = (u - a.npix[n * 2]) * ((float)a.W * 0.5f), fy = (v - a.npix[n * 2 + 1]) * ((float)a.H * 0.5f);
  G.cx = a.pix[n * 2] + fx;
  G.cy = a.pix[n * 2 + 1] + fy;
  G.valid = G.cx >= 0.0f && G.cx < (float)a.W && G.cy >= 0.0f && G.cy < (float)a.H;
  // warp_pixel normalises by (size - 1) / 2, grid_sample(align_corners=True) undoes it
  const float gx = G.cx / ((float)(a.W - 1) * 0.5f) - 1.0f, gy = G.cy / ((float)(a.H - 1) * 0.5f) - 1.0f;
  G.ix = clip_border((gx + 1.0f) * 0.5f * (float)(a.W - 1), a.W, &G.mx);
  G.iy = clip_border((gy + 1.0f) * 0.5f * (float)(a.H - 1), a.H, &G.my);
  G.x0 = (int)floorf(G.ix);
  G.y0 = (int)floorf(G.iy);
  return G;
}

__device__ __forceinline__ float tap(const float* img, int H, int W, int y, int x) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? img[(int64_t)y * W + x] : 0.0f;
}

// ws: [2t] sum |warped - gt| over valid rays, [2t+1] valid count, [2T] block counter; one thread per (ray, frame)
__global__ void __launch_bounds__(kLossThreads) flow_rgb_fwd_kernel(const FlowRgbArgs a, float* __restrict__ flow_pred,
                                                                   float* __restrict__ ws, float* __restrict__ loss) {
  __shared__ float sh[2 * 8];
  const int t = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc[2] = {0.0f, 0.0f};
  if (n < a.N) {
    const WarpGeom G = flow_geom(a, n, t);
    if (flow_pred) {
      flow_pred[((int64_t)t * a.N + n) * 2] = G.cx - a.pix[n * 2];
      flow_pred[((int64_t)t * a.N + n) * 2 + 1] = G.cy - a.pix[n * 2 + 1];
    }
    if (G.valid) {
      const float tx = G.ix - (float)G.x0, ty = G.iy - (float)G.y0;
      const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty), w10 = (1.0f - tx) * ty, w11 = tx * ty;
      for (int c = 0; c < 3; ++c) {
        const float* img = a.ref + ((int64_t)t * 3 + c) * a.H * a.W;
        const float v = w00 * tap(img, a.H, a.W, G.y0, G.x0) + w01 * tap(img, a.H, a.W, G.y0, G.x0 + 1) +
                        w10 * tap(img, a.H, a.W, G.y0 + 1, G.x0) + w11 * tap(img, a.H, a.W, G.y0 + 1, G.x0 + 1);
        acc[0] += fabsf(v - a.rgb_gt[n * 3 + c]);
      }
      acc[1] = 1.0f;
    }
  }
  block_sum<2>(acc, sh);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    if (acc[0] != 0.0f) atomicAdd(ws + 2 * t, acc[0]);
    if (acc[1] != 0.0f) atomicAdd(ws + 2 * t + 1, acc[1]);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(ws + 2 * a.T), 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* v = ws;
    float l = 0.0f;
    for (int k = 0; k < a.T; ++k) l += v[2 * k] / (v[2 * k + 1] + 1e-10f);
    loss[0] = l / 3.0f;                                  // train.py:517 divides by 3 whatever the number of frames
  }
}

// one thread per ray, frames in the inner loop (d_wp sums over frames); d_w2c rows 0..2 are block-reduced per frame
__global__ void __launch_bounds__(kLossThreads) flow_rgb_bwd_kernel(const FlowRgbArgs a, const float* __restrict__ ws,
                                                                   const float* __restrict__ g_ptr, float4* __restrict__ d_wp,
                                                                   float* __restrict__ d_w2c) {
  __shared__ float sh[12 * 8];
  const float g = (g_ptr ? *g_ptr : 1.0f) / 3.0f;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4 dwp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (int t = 0; t < a.T; ++t) {
    float dM[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (n < a.N) {
      const WarpGeom G = flow_geom(a, n, t);
      if (G.valid) {
        const float cg = g / (ws[2 * t + 1] + 1e-10f);
        const float tx = G.ix - (float)G.x0, ty = G.iy - (float)G.y0;
        const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty), w10 = (1.0f - tx) * ty, w11 = tx * ty;
        float gix = 0.0f, giy = 0.0f;
        for (int c = 0; c < 3; ++c) {
          const float* img = a.ref + ((int64_t)t * 3 + c) * a.H * a.W;
          const float v00 = tap(img, a.H, a.W, G.y0, G.x0), v01 = tap(img, a.H, a.W, G.y0, G.x0 + 1),
                      v10 = tap(img, a.H, a.W, G.y0 + 1, G.x0), v11 = tap(img, a.H, a.W, G.y0 + 1, G.x0 + 1);
          const float v = w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11;
          const float go = cg * signf_(v - a.rgb_gt[n * 3 + c]);
          gix += go * ((v01 - v00) * (1.0f - ty) + (v11 - v10) * ty);
          giy += go * ((v10 - v00) * (1.0f - tx) + (v11 - v01) * tx);
        }
        // d ix / d cx = 1 (normalise and un-normalise cancel), times the clamp's derivative
        const float du = gix * G.mx * ((float)a.W * 0.5f), dv = giy * G.my * ((float)a.H * 0.5f);
        const float iz = 1.0f / G.q[2];
        const float dq[3] = {du * iz, dv * iz, -(du * G.q[0] + dv * G.q[1]) * iz * iz};
        const float* K = a.KS + t * 9;
        const float* M = a.w2c + t * 16;
        const float4 w = a.wp[n];
        float dm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) dm[k] = K[k] * dq[0] + K[3 + k] * dq[1] + K[6 + k] * dq[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          dM[r * 4] = dm[r] * w.x; dM[r * 4 + 1] = dm[r] * w.y; dM[r * 4 + 2] = dm[r] * w.z; dM[r * 4 + 3] = dm[r] * w.w;
        }
        dwp.x += M[0] * dm[0] + M[4] * dm[1] + M[8] * dm[2];
        dwp.y += M[1] * dm[0] + M[5] * dm[1] + M[9] * dm[2];
        dwp.z += M[2] * dm[0] + M[6] * dm[1] + M[10] * dm[2];
        dwp.w += M[3] * dm[0] + M[7] * dm[1] + M[11] * dm[2];
      }
    }
    if (d_w2c) {
      block_sum<12>(dM, sh);
      if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < 12; ++k)
          if (dM[k] != 0.0f) atomicAdd(d_w2c + t * 16 + k, dM[k]);
    }
  }
  if (n < a.N && d_wp) d_wp[n] = dwp;
}


// ------------------------------------------------------------------------------------------------ depth-patch smoothness
// model/losses.py:7-38 on ps x ps patches of depth_pred (train.py:519-525): four neighbour differences per patch
//   D1 (r,c)-(r,c+1)   D2 (r,c)-(r+1,c)   D3 (r,c)-(r+1,c+1)   D4 (r+1,c)-(r,c+1)
// smooth = mean_k mean|D_k|, edge = mean_k mean(w_k |D_k|) with the bilateral weight w_k = exp(-sum_c |rgb diff_k| / gamma).
// One thread per patch.  ws: [0..3] sum |D_k|, [4..7] sum w_k |D_k|, [8] block counter.
constexpr int kMaxPatch = 8;
struct PatchArgs {
  const float* depth;      // [n x ps x ps]
  const float* rgb;        // [n x ps x ps x 3] (null: no edge-aware term)
  int64_t n; int ps; float inv_gamma, w_edge, w_smooth;
};
__device__ __forceinline__ void patch_pair(int k, int r, int c, int ps, int* i0, int* i1) {
  // element indices (inside the patch) of the minuend / subtrahend of difference k at position (r, c)
  switch (k) {
    case 0: *i0 = r * ps + c; *i1 = r * ps + c + 1; break;
    case 1: *i0 = r * ps + c; *i1 = (r + 1) * ps + c; break;
    case 2: *i0 = r * ps + c; *i1 = (r + 1) * ps + c + 1; break;
    default: *i0 = (r + 1) * ps + c; *i1 = r * ps + c + 1; break;
  }
}
__device__ __forceinline__ float bilateral(const float* rgb, int i0, int i1, float inv_gamma) {
  const float d = fabsf(rgb[i0 * 3] - rgb[i1 * 3]) + fabsf(rgb[i0 * 3 + 1] - rgb[i1 * 3 + 1]) + fabsf(rgb[i0 * 3 + 2] - rgb[i1 * 3 + 2]);
  return __expf(-d * inv_gamma);
}
__global__ void __launch_bounds__(kLossThreads) patch_smooth_fwd_kernel(const PatchArgs a, float* __restrict__ ws,
                                                                       float* __restrict__ losses) {
  __shared__ float sh[8 * 8];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int ps = a.ps, pp = ps * ps;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.n; n += (int64_t)gridDim.x * blockDim.x) {
    const float* d = a.depth + n * pp;
    const float* rgb = a.rgb ? a.rgb + n * pp * 3 : nullptr;
    for (int k = 0; k < 4; ++k) {
      const int R = (k == 0) ? ps : ps - 1, Cn = (k == 1) ? ps : ps - 1;
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < Cn; ++c) {
          int i0, i1;
          patch_pair(k, r, c, ps, &i0, &i1);
          const float ad = fabsf(d[i0] - d[i1]);
          acc[k] += ad;
          if (rgb) acc[4 + k] += bilateral(rgb, i0, i1, a.inv_gamma) * ad;
        }
    }
  }
  block_sum<8>(acc, sh);
  __shared__ bool last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (acc[k] != 0.0f) atomicAdd(ws + k, acc[k]);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(ws + 8), 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* v = ws;
    const float c_a = 1.0f / ((float)a.n * (float)(ps * (ps - 1))), c_b = 1.0f / ((float)a.n * (float)((ps - 1) * (ps - 1)));
    const float smooth = 0.25f * ((v[0] + v[1]) * c_a + (v[2] + v[3]) * c_b);
    const float edge = 0.25f * ((v[4] + v[5]) * c_a + (v[6] + v[7]) * c_b);
    losses[0] = a.w_edge * edge + a.w_smooth * smooth;
    losses[1] = edge; losses[2] = smooth;
  }
}
__global__ void __launch_bounds__(kLossThreads) patch_smooth_bwd_kernel(const PatchArgs a, const float* __restrict__ g_ptr,
                                                                       float* __restrict__ d_depth) {
  const float g = g_ptr ? *g_ptr : 1.0f;
  const int ps = a.ps, pp = ps * ps;
  const float c_a = 0.25f * g / ((float)a.n * (float)(ps * (ps - 1))), c_b = 0.25f * g / ((float)a.n * (float)((ps - 1) * (ps - 1)));
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < a.n; n += (int64_t)gridDim.x * blockDim.x) {
    const float* d = a.depth + n * pp;
    const float* rgb = a.rgb ? a.rgb + n * pp * 3 : nullptr;
    float* o = d_depth + n * pp;                 // each patch is owned by one thread: plain read-modify-write
    for (int i = 0; i < pp; ++i) o[i] = 0.0f;
    for (int k = 0; k < 4; ++k) {
      const int R = (k == 0) ? ps : ps - 1, Cn = (k == 1) ? ps : ps - 1;
      const float ck = k < 2 ? c_a : c_b;
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < Cn; ++c) {
          int i0, i1;
          patch_pair(k, r, c, ps, &i0, &i1);
          const float w = a.w_smooth + (rgb ? a.w_edge * bilateral(rgb, i0, i1, a.inv_gamma) : 0.0f);
          const float t = ck * w * signf_(d[i0] - d[i1]);
          o[i0] += t; o[i1] -= t;
        }
    }
  }
}

static inline unsigned loss_grid(int64_t n) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kLossThreads), 148 * 8));
}

}  // namespace
}  // namespace cope

using namespace cope;

extern "C" {

int cope_step_losses_fwd(const float* color, const float* rgb_gt, const float* grad4, const float* pts4, const float* weights,
                         const float* motion, const float* w_sum_global, int64_t N, int64_t P, float w_rgb, float w_eik,
                         float w_flow, float* losses, float* coef, float* ws, cope_stream_t s) {
  COPE_REQUIRE(N >= 0 && P >= 0, "step_losses: negative sizes");
  COPE_REQUIRE(!motion || (pts4 && weights), "step_losses: the SDF-flow term needs pts4 and weights");
  cudaMemsetAsync(ws, 0, 8 * sizeof(float), as_stream(s));
  StepLossArgs a{color, rgb_gt, reinterpret_cast<const float4*>(grad4), reinterpret_cast<const float4*>(pts4), weights, motion,
                 w_sum_global, N, P, w_rgb, w_eik, w_flow};
  step_losses_fwd_kernel<<<loss_grid(std::max(P, N * 3)), kLossThreads, 0, as_stream(s)>>>(a, ws, losses, coef);
  COPE_CHECK_LAUNCH("step_losses_fwd");
  return 0;
}

int cope_step_losses_bwd(const float* color, const float* rgb_gt, const float* grad4, const float* pts4, const float* weights,
                         const float* motion, int64_t N, int64_t P, const float* coef, const float* g, float* d_color,
                         float* d_grad4, float* d_pts4, float* d_motion, cope_stream_t s) {
  COPE_REQUIRE(!motion || (pts4 && weights), "step_losses: the SDF-flow term needs pts4 and weights");
  if (N <= 0 && P <= 0) return 0;
  StepLossArgs a{color, rgb_gt, reinterpret_cast<const float4*>(grad4), reinterpret_cast<const float4*>(pts4), weights, motion,
                 nullptr, N, P, 0.0f, 0.0f, 0.0f};
  step_losses_bwd_kernel<<<loss_grid(std::max(P, N * 3)), kLossThreads, 0, as_stream(s)>>>(
      a, coef, g, d_color, reinterpret_cast<float4*>(d_grad4), reinterpret_cast<float4*>(d_pts4), d_motion);
  COPE_CHECK_LAUNCH("step_losses_bwd");
  return 0;
}

int cope_weighted_points_fwd(const float* weights, const float* pts4, int64_t N, int S, float* wp, cope_stream_t s) {
  if (N <= 0) return 0;
  weighted_points_fwd_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, as_stream(s)>>>(weights, reinterpret_cast<const float4*>(pts4), N, S,
                                                                                reinterpret_cast<float4*>(wp));
  COPE_CHECK_LAUNCH("weighted_points_fwd");
  return 0;
}

int cope_weighted_points_bwd(const float* weights, const float* pts4, const float* d_wp, int64_t N, int S, float* d_weights,
                             float* d_pts4, cope_stream_t s) {
  if (N <= 0) return 0;
  const int64_t P = N * S;
  weighted_points_bwd_kernel<<<(unsigned)ceil_div(P, 256), 256, 0, as_stream(s)>>>(
      weights, reinterpret_cast<const float4*>(pts4), reinterpret_cast<const float4*>(d_wp), P, S, d_weights,
      reinterpret_cast<float4*>(d_pts4));
  COPE_CHECK_LAUNCH("weighted_points_bwd");
  return 0;
}

int cope_flow_rgb_fwd(const float* wp, const float* w2c, const float* KS, const float* norm_pix, const float* pix,
                      const float* ref_imgs, const float* rgb_gt, int64_t N, int T, int H, int W, float* flow_pred, float* loss,
                      float* ws, cope_stream_t s) {
  COPE_REQUIRE(T >= 1 && T <= 16 && H >= 2 && W >= 2, "flow_rgb: T=%d H=%d W=%d out of range", T, H, W);
  cudaMemsetAsync(ws, 0, (2 * T + 1) * sizeof(float), as_stream(s));
  if (N <= 0) { cudaMemsetAsync(loss, 0, sizeof(float), as_stream(s)); return 0; }
  FlowRgbArgs a{reinterpret_cast<const float4*>(wp), w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, N, T, H, W};
  flow_rgb_fwd_kernel<<<dim3((unsigned)ceil_div(N, kLossThreads), T), kLossThreads, 0, as_stream(s)>>>(a, flow_pred, ws, loss);
  COPE_CHECK_LAUNCH("flow_rgb_fwd");
  return 0;
}

int cope_flow_rgb_bwd(const float* wp, const float* w2c, const float* KS, const float* norm_pix, const float* pix,
                      const float* ref_imgs, const float* rgb_gt, int64_t N, int T, int H, int W, const float* ws, const float* g,
                      float* d_wp, float* d_w2c, cope_stream_t s) {
  COPE_REQUIRE(T >= 1 && T <= 16 && H >= 2 && W >= 2, "flow_rgb: T=%d H=%d W=%d out of range", T, H, W);
  if (N <= 0) return 0;
  FlowRgbArgs a{reinterpret_cast<const float4*>(wp), w2c, KS, norm_pix, pix, ref_imgs, rgb_gt, N, T, H, W};
  flow_rgb_bwd_kernel<<<(unsigned)ceil_div(N, kLossThreads), kLossThreads, 0, as_stream(s)>>>(a, ws, g, reinterpret_cast<float4*>(d_wp),
                                                                                              d_w2c);
  COPE_CHECK_LAUNCH("flow_rgb_bwd");
  return 0;
}

int cope_patch_smooth_fwd(const float* depth, const float* rgb, int64_t n_patches, int ps, float gamma, float w_edge,
                          float w_smooth, float* losses, float* ws, cope_stream_t s) {
  COPE_REQUIRE(ps >= 2 && ps <= kMaxPatch, "patch_smooth: patch size %d outside [2, %d]", ps, kMaxPatch);
  COPE_REQUIRE(gamma > 0.0f, "patch_smooth: gamma must be positive");
  cudaMemsetAsync(ws, 0, 12 * sizeof(float), as_stream(s));
  if (n_patches <= 0) { cudaMemsetAsync(losses, 0, 3 * sizeof(float), as_stream(s)); return 0; }
  PatchArgs a{depth, rgb, n_patches, ps, 1.0f / gamma, w_edge, w_smooth};
  patch_smooth_fwd_kernel<<<loss_grid(n_patches), kLossThreads, 0, as_stream(s)>>>(a, ws, losses);
  COPE_CHECK_LAUNCH("patch_smooth_fwd");
  return 0;
}

int cope_patch_smooth_bwd(const float* depth, const float* rgb, int64_t n_patches, int ps, float gamma, float w_edge,
                          float w_smooth, const float* g, float* d_depth, cope_stream_t s) {
  COPE_REQUIRE(ps >= 2 && ps <= kMaxPatch, "patch_smooth: patch size %d outside [2, %d]", ps, kMaxPatch);
  if (n_patches <= 0) return 0;
  PatchArgs a{depth, rgb, n_patches, ps, 1.0f / gamma, w_edge, w_smooth};
  patch_smooth_bwd_kernel<<<loss_grid(n_patches), kLossThreads, 0, as_stream(s)>>>(a, g, d_depth);
  COPE_CHECK_LAUNCH("patch_smooth_bwd");
  return 0;
}

}  // extern "C"
