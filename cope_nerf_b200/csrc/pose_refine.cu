// Photometric relative-pose refinement between the two training stages:
// utils_poses/pose_refinement.py:34-61 (compute_loss_and_warp_image) as ONE forward and ONE backward launch.
//
//   per pixel (u, v) of image b, normalised to [-1, 1] as pose_refinement.py:88-96 builds `uv`:
//     xyz  = K_b^-1 (u d, v d, d)                     d = depths[b, row, col]
//     xyz' = R_b xyz + T_b                            relative_poses[b] = [R | T]
//     q    = K_b xyz' ;  (u', v') = q.xy / q.z
//     valid = |u'| <= 1 and |v'| <= 1
//     warped = grid_sample(next_images[b], (u', v'), bilinear, border, align_corners=True)     (train.py:235-244, normalize_pix=False)
//   loss = sum_{b, c, pixel} |warped - images| valid / sum_{b, pixel} valid
//
// The only trainable input is the pose (PoseRetriever r, t of the pair): the backward returns d loss / d relative_poses [B, 3x4],
// block-reduced per image.  HBM-trivial (one read of two frames + a depth map); the point is launch count: the reference spends
// ~40 ATen launches (3 batched inverses / matmuls, grid_sample, masks) per direction per batch.
#include <algorithm>

#include "common.cuh"

namespace cope {
namespace {

constexpr int kThreads = 256;

struct RefineArgs {
  const float* img;        // [B x 3 x H x W] target frames
  const float* next;       // [B x 3 x H x W] frames that are warped
  const float* depth;      // [B x H x W]
  const float* K;          // [B x 9]
  const float* pose;       // [B x 16] row-major 4 x 4
  int B, H, W;
};

__device__ __forceinline__ float sgn(float x) { return (x > 0.0f) - (x < 0.0f); }

// general 3 x 3 inverse (adjugate); K is upper-triangular in practice but nothing here relies on it
__device__ __forceinline__ void inv3(const float* m, float* o) {
  const float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  const float A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const float inv = 1.0f / (a * A + b * Bc + c * C);
  o[0] = A * inv; o[1] = -(b * i - c * h) * inv; o[2] = (b * f - c * e) * inv;
  o[3] = Bc * inv; o[4] = (a * i - c * g) * inv; o[5] = -(a * f - c * d) * inv;
  o[6] = C * inv; o[7] = -(a * h - b * g) * inv; o[8] = (a * e - b * d) * inv;
}

struct Geom {
  float x[3];              // xyz in the source camera
  float q[3];              // K (R xyz + T)
  float up, vp;            // projected normalised coordinate
  bool valid;
  float ix, iy, mx, my;    // clamped pixel coordinate in the warped frame and the clamp's derivative
  int x0, y0;
};

__device__ __forceinline__ float clip_border(float v, int size, float* mult) {      // grid_sample padding_mode='border'
  if (v <= 0.0f) { *mult = 0.0f; return 0.0f; }
  const float mx = (float)(size - 1);
  if (v >= mx) { *mult = 0.0f; return mx; }
  *mult = 1.0f;
  return v;
}

__device__ __forceinline__ Geom geom(const RefineArgs& a, const float* Ki, const float* K, const float* M, int b, int row, int col) {
  Geom G;
  const float u = (float)col / ((float)(a.W - 1) * 0.5f) - 1.0f, v = (float)row / ((float)(a.H - 1) * 0.5f) - 1.0f;
  const float d = a.depth[((int64_t)b * a.H + row) * a.W + col];
  const float p0 = u * d, p1 = v * d, p2 = d;
#pragma unroll
  for (int r = 0; r < 3; ++r) G.x[r] = Ki[r * 3] * p0 + Ki[r * 3 + 1] * p1 + Ki[r * 3 + 2] * p2;
  float t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) t[r] = M[r * 4] * G.x[0] + M[r * 4 + 1] * G.x[1] + M[r * 4 + 2] * G.x[2] + M[r * 4 + 3];
#pragma unroll
  for (int r = 0; r < 3; ++r) G.q[r] = K[r * 3] * t[0] + K[r * 3 + 1] * t[1] + K[r * 3 + 2] * t[2];
  G.up = G.q[0] / G.q[2];
  G.vp = G.q[1] / G.q[2];
  G.valid = G.up >= -1.0f && G.up <= 1.0f && G.vp >= -1.0f && G.vp <= 1.0f;
  G.ix = clip_border((G.up + 1.0f) * 0.5f * (float)(a.W - 1), a.W, &G.mx);
  G.iy = clip_border((G.vp + 1.0f) * 0.5f * (float)(a.H - 1), a.H, &G.my);
  G.x0 = (int)floorf(G.ix);
  G.y0 = (int)floorf(G.iy);
  return G;
}

__device__ __forceinline__ float tap(const float* img, int H, int W, int y, int x) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? img[(int64_t)y * W + x] : 0.0f;
}

template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* sh) {
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

// grid: (pixel blocks, B).  ws: [0] sum |warped - image| valid, [1] sum valid, [2] block counter
__global__ void __launch_bounds__(kThreads) pose_refine_fwd_kernel(const RefineArgs a, float* __restrict__ warped,
                                                                  float* __restrict__ ws, float* __restrict__ loss) {
  __shared__ float sK[9], sKi[9], sM[16], sh[2 * 8];
  const int b = blockIdx.y;
  if (threadIdx.x < 9) sK[threadIdx.x] = a.K[b * 9 + threadIdx.x];
  if (threadIdx.x < 16) sM[threadIdx.x] = a.pose[b * 16 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) inv3(sK, sKi);
  __syncthreads();
  const int HW = a.H * a.W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[2] = {0.0f, 0.0f};
  if (p < HW) {
    const int row = p / a.W, col = p - row * a.W;
    const Geom G = geom(a, sKi, sK, sM, b, row, col);
    const float tx = G.ix - (float)G.x0, ty = G.iy - (float)G.y0;
    const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty), w10 = (1.0f - tx) * ty, w11 = tx * ty;
    for (int c = 0; c < 3; ++c) {
      const float* src = a.next + ((int64_t)b * 3 + c) * HW;
      const float v = w00 * tap(src, a.H, a.W, G.y0, G.x0) + w01 * tap(src, a.H, a.W, G.y0, G.x0 + 1) +
                      w10 * tap(src, a.H, a.W, G.y0 + 1, G.x0) + w11 * tap(src, a.H, a.W, G.y0 + 1, G.x0 + 1);
      if (warped) warped[((int64_t)b * 3 + c) * HW + p] = v;
      if (G.valid) acc[0] += fabsf(v - a.img[((int64_t)b * 3 + c) * HW + p]);
    }
    if (G.valid) acc[1] = 1.0f;
  }
  block_sum<2>(acc, sh);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    if (acc[0] != 0.0f) atomicAdd(ws, acc[0]);
    if (acc[1] != 0.0f) atomicAdd(ws + 1, acc[1]);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned*>(ws + 2), 1u) == gridDim.x * gridDim.y - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const volatile float* v = ws;
    loss[0] = v[0] / v[1];                      // torch.sum(valid_mask) without an epsilon (pose_refinement.py:59)
  }
}

// d loss / d pose[b] rows 0..2 (12 numbers), block-reduced; g = upstream gradient of the loss (device scalar)
__global__ void __launch_bounds__(kThreads) pose_refine_bwd_kernel(const RefineArgs a, const float* __restrict__ ws,
                                                                  const float* __restrict__ g_ptr, float* __restrict__ d_pose) {
  __shared__ float sK[9], sKi[9], sM[16], sh[12 * 8];
  const int b = blockIdx.y;
  if (threadIdx.x < 9) sK[threadIdx.x] = a.K[b * 9 + threadIdx.x];
  if (threadIdx.x < 16) sM[threadIdx.x] = a.pose[b * 16 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) inv3(sK, sKi);
  __syncthreads();
  const int HW = a.H * a.W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const float cg = (g_ptr ? *g_ptr : 1.0f) / ws[1];
  float dM[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (p < HW) {
    const int row = p / a.W, col = p - row * a.W;
    const Geom G = geom(a, sKi, sK, sM, b, row, col);
    if (G.valid) {
      const float tx = G.ix - (float)G.x0, ty = G.iy - (float)G.y0;
      const float w00 = (1.0f - tx) * (1.0f - ty), w01 = tx * (1.0f - ty), w10 = (1.0f - tx) * ty, w11 = tx * ty;
      float gix = 0.0f, giy = 0.0f;
      for (int c = 0; c < 3; ++c) {
        const float* src = a.next + ((int64_t)b * 3 + c) * HW;
        const float v00 = tap(src, a.H, a.W, G.y0, G.x0), v01 = tap(src, a.H, a.W, G.y0, G.x0 + 1),
                    v10 = tap(src, a.H, a.W, G.y0 + 1, G.x0), v11 = tap(src, a.H, a.W, G.y0 + 1, G.x0 + 1);
        const float v = w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11;
        const float go = cg * sgn(v - a.img[((int64_t)b * 3 + c) * HW + p]);
        gix += go * ((v01 - v00) * (1.0f - ty) + (v11 - v10) * ty);
        giy += go * ((v10 - v00) * (1.0f - tx) + (v11 - v01) * tx);
      }
      // ix = (u' + 1) / 2 * (W - 1), clamped
      const float du = gix * G.mx * 0.5f * (float)(a.W - 1), dv = giy * G.my * 0.5f * (float)(a.H - 1);
      const float iz = 1.0f / G.q[2];
      const float dq[3] = {du * iz, dv * iz, -(du * G.q[0] + dv * G.q[1]) * iz * iz};
      float dt[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) dt[k] = sK[k] * dq[0] + sK[3 + k] * dq[1] + sK[6 + k] * dq[2];     // K^T dq
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        dM[r * 4] = dt[r] * G.x[0]; dM[r * 4 + 1] = dt[r] * G.x[1]; dM[r * 4 + 2] = dt[r] * G.x[2]; dM[r * 4 + 3] = dt[r];
      }
    }
  }
  block_sum<12>(dM, sh);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 12; ++k)
      if (dM[k] != 0.0f) atomicAdd(d_pose + b * 16 + k, dM[k]);
}

}  // namespace
}  // namespace cope

using namespace cope;

extern "C" {

int cope_pose_refine_fwd(const float* images, const float* next_images, const float* depths, const float* K, const float* poses,
                         int B, int H, int W, float* warped, float* loss, float* ws, cope_stream_t s) {
  COPE_REQUIRE(B >= 1 && H >= 2 && W >= 2, "pose_refine: B=%d H=%d W=%d out of range", B, H, W);
  COPE_REQUIRE(images && next_images && depths && K && poses && loss && ws, "pose_refine: null argument");
  cudaMemsetAsync(ws, 0, 4 * sizeof(float), as_stream(s));
  RefineArgs a{images, next_images, depths, K, poses, B, H, W};
  pose_refine_fwd_kernel<<<dim3((unsigned)ceil_div((int64_t)H * W, kThreads), B), kThreads, 0, as_stream(s)>>>(a, warped, ws, loss);
  COPE_CHECK_LAUNCH("pose_refine_fwd");
  return 0;
}

int cope_pose_refine_bwd(const float* images, const float* next_images, const float* depths, const float* K, const float* poses,
                         int B, int H, int W, const float* ws, const float* g, float* d_poses, cope_stream_t s) {
  COPE_REQUIRE(B >= 1 && H >= 2 && W >= 2, "pose_refine: B=%d H=%d W=%d out of range", B, H, W);
  COPE_REQUIRE(images && next_images && depths && K && poses && ws && d_poses, "pose_refine: null argument");
  RefineArgs a{images, next_images, depths, K, poses, B, H, W};
  pose_refine_bwd_kernel<<<dim3((unsigned)ceil_div((int64_t)H * W, kThreads), B), kThreads, 0, as_stream(s)>>>(a, ws, g, d_poses);
  COPE_CHECK_LAUNCH("pose_refine_bwd");
  return 0;
}

}  // extern "C"
