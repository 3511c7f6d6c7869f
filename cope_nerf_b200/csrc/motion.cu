// Continuous pose model (MotionNetwork, model/neus_fields.py:142-183): integrate the predicted angular / linear velocities
// over the sub-steps of every consecutive frame pair and chain the relative poses into world -> camera maps.
//
//   per pair f, sub-step k (compute_consecutive_relative_pose, :152-160):
//       R_k = Rx(w_x dt) Ry(w_y dt) Rz(w_z dt)   (euler_angles_to_matrix(.., 'XYZ'), utils_poses/pose_pytorch3d.py:8-19)
//       T <- R_k T + v dt ;  R <- R R_k           rel_f = [R T; 0 1]
//   chain (compute_w2c_mappings, :172-183):  w2c_0 = I,  w2c_{i+1} = rel_i w2c_i
//
// The reference runs this as a Python double loop (thousands of tiny launches per training step); here the pairs are
// independent threads (the sub-step recursion is 10 steps long) and the chain is one short sequential product.
// Everything is 3x3 / 4x4 fp32 arithmetic: latency-, not throughput-bound.
#include "common.cuh"

namespace cope {
namespace {

constexpr int kMaxSub = 64;

struct M3 { float m[9]; };
__device__ __forceinline__ M3 mul3(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return c;
}
__device__ __forceinline__ M3 mul3_at(const M3& a, const M3& b) {   // a^T b
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[i * 3 + j] = a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j] + a.m[6 + i] * b.m[6 + j];
  return c;
}
__device__ __forceinline__ M3 mul3_bt(const M3& a, const M3& b) {   // a b^T
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[i * 3 + j] = a.m[i * 3] * b.m[j * 3] + a.m[i * 3 + 1] * b.m[j * 3 + 1] + a.m[i * 3 + 2] * b.m[j * 3 + 2];
  return c;
}
// R = Rx(a) Ry(b) Rz(c)
__device__ __forceinline__ M3 euler_xyz(float a, float b, float c, float (&sc)[6]) {
  float sa, ca, sb, cb, sg, cg;
  sincosf(a, &sa, &ca); sincosf(b, &sb, &cb); sincosf(c, &sg, &cg);
  sc[0] = sa; sc[1] = ca; sc[2] = sb; sc[3] = cb; sc[4] = sg; sc[5] = cg;
  M3 r;
  r.m[0] = cb * cg;                 r.m[1] = -cb * sg;                r.m[2] = sb;
  r.m[3] = ca * sg + sa * sb * cg;  r.m[4] = ca * cg - sa * sb * sg;  r.m[5] = -sa * cb;
  r.m[6] = sa * sg - ca * sb * cg;  r.m[7] = sa * cg + ca * sb * sg;  r.m[8] = ca * cb;
  return r;
}
// d(angles) from dR for R = Rx(a) Ry(b) Rz(c)
__device__ __forceinline__ void euler_xyz_bwd(const float (&sc)[6], const M3& g, float& da, float& db, float& dc) {
  const float sa = sc[0], ca = sc[1], sb = sc[2], cb = sc[3], sg = sc[4], cg = sc[5];
  da = g.m[3] * (-sa * sg + ca * sb * cg) + g.m[4] * (-sa * cg - ca * sb * sg) + g.m[5] * (-ca * cb) +
       g.m[6] * (ca * sg + sa * sb * cg) + g.m[7] * (ca * cg - sa * sb * sg) + g.m[8] * (-sa * cb);
  db = g.m[0] * (-sb * cg) + g.m[1] * (sb * sg) + g.m[2] * cb + g.m[3] * (sa * cb * cg) + g.m[4] * (-sa * cb * sg) +
       g.m[5] * (sa * sb) + g.m[6] * (-ca * cb * cg) + g.m[7] * (ca * cb * sg) + g.m[8] * (-ca * sb);
  dc = g.m[0] * (-cb * sg) + g.m[1] * (-cb * cg) + g.m[3] * (ca * cg - sa * sb * sg) + g.m[4] * (-ca * sg - sa * sb * cg) +
       g.m[6] * (sa * cg + ca * sb * sg) + g.m[7] * (-sa * sg + ca * sb * cg);
}

// one thread per frame pair
__global__ void pose_integrate_fwd_kernel(const float* __restrict__ wv, const float* __restrict__ dt_p, int F, int n_sub,
                                          float* __restrict__ rel) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float dt = dt_p[f];          // every pair integrates with its own time interval (:148)
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  float T[3] = {0.0f, 0.0f, 0.0f};
  for (int k = 0; k < n_sub; ++k) {
    const float* q = wv + ((int64_t)f * n_sub + k) * 6;
    float sc[6];
    const M3 Rk = euler_xyz(q[0] * dt, q[1] * dt, q[2] * dt, sc);
    const float t0 = Rk.m[0] * T[0] + Rk.m[1] * T[1] + Rk.m[2] * T[2] + q[3] * dt;
    const float t1 = Rk.m[3] * T[0] + Rk.m[4] * T[1] + Rk.m[5] * T[2] + q[4] * dt;
    const float t2 = Rk.m[6] * T[0] + Rk.m[7] * T[1] + Rk.m[8] * T[2] + q[5] * dt;
    T[0] = t0; T[1] = t1; T[2] = t2;
    R = mul3(R, Rk);
  }
  float* o = rel + (int64_t)f * 16;
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i * 4] = R.m[i * 3]; o[i * 4 + 1] = R.m[i * 3 + 1]; o[i * 4 + 2] = R.m[i * 3 + 2]; o[i * 4 + 3] = T[i]; }
  o[12] = 0.0f; o[13] = 0.0f; o[14] = 0.0f; o[15] = 1.0f;
}

// reverse of the recursion; d_wv [F * n_sub, 6] and d_dt [F] OVERWRITTEN
__global__ void pose_integrate_bwd_kernel(const float* __restrict__ wv, const float* __restrict__ dt_p, int F, int n_sub,
                                          const float* __restrict__ d_rel, float* __restrict__ d_wv, float* __restrict__ d_dt) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float dt = dt_p[f];          // every pair integrates with its own time interval (:148)
  // forward again, keeping the state BEFORE every sub-step (n_sub <= kMaxSub)
  M3 Rs[kMaxSub];
  float Ts[kMaxSub][3];
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  float T[3] = {0.0f, 0.0f, 0.0f};
  for (int k = 0; k < n_sub; ++k) {
    Rs[k] = R; Ts[k][0] = T[0]; Ts[k][1] = T[1]; Ts[k][2] = T[2];
    const float* q = wv + ((int64_t)f * n_sub + k) * 6;
    float sc[6];
    const M3 Rk = euler_xyz(q[0] * dt, q[1] * dt, q[2] * dt, sc);
    const float t0 = Rk.m[0] * T[0] + Rk.m[1] * T[1] + Rk.m[2] * T[2] + q[3] * dt;
    const float t1 = Rk.m[3] * T[0] + Rk.m[4] * T[1] + Rk.m[5] * T[2] + q[4] * dt;
    const float t2 = Rk.m[6] * T[0] + Rk.m[7] * T[1] + Rk.m[8] * T[2] + q[5] * dt;
    T[0] = t0; T[1] = t1; T[2] = t2;
    R = mul3(R, Rk);
  }
  const float* g = d_rel + (int64_t)f * 16;
  M3 gR;
  float gT[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { gR.m[i * 3] = g[i * 4]; gR.m[i * 3 + 1] = g[i * 4 + 1]; gR.m[i * 3 + 2] = g[i * 4 + 2]; gT[i] = g[i * 4 + 3]; }
  float ddt = 0.0f;
  for (int k = n_sub - 1; k >= 0; --k) {
    const float* q = wv + ((int64_t)f * n_sub + k) * 6;
    float sc[6];
    const M3 Rk = euler_xyz(q[0] * dt, q[1] * dt, q[2] * dt, sc);
    // R_out = R_in Rk  ->  dRk += R_in^T gR ; gR_in = gR Rk^T
    M3 gRk = mul3_at(Rs[k], gR);
    gR = mul3_bt(gR, Rk);
    // T_out = Rk T_in + v dt  ->  dRk += gT T_in^T ; d(v dt) = gT ; gT_in = Rk^T gT
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) gRk.m[i * 3 + j] += gT[i] * Ts[k][j];
    float da, db, dc;
    euler_xyz_bwd(sc, gRk, da, db, dc);
    float* o = d_wv + ((int64_t)f * n_sub + k) * 6;
    o[0] = da * dt; o[1] = db * dt; o[2] = dc * dt;
    o[3] = gT[0] * dt; o[4] = gT[1] * dt; o[5] = gT[2] * dt;
    ddt += da * q[0] + db * q[1] + dc * q[2] + gT[0] * q[3] + gT[1] * q[4] + gT[2] * q[5];
    const float n0 = Rk.m[0] * gT[0] + Rk.m[3] * gT[1] + Rk.m[6] * gT[2];
    const float n1 = Rk.m[1] * gT[0] + Rk.m[4] * gT[1] + Rk.m[7] * gT[2];
    const float n2 = Rk.m[2] * gT[0] + Rk.m[5] * gT[1] + Rk.m[8] * gT[2];
    gT[0] = n0; gT[1] = n1; gT[2] = n2;
  }
  if (d_dt) d_dt[f] = ddt;
}

// w2c_0 = I, w2c_{i+1} = rel_i w2c_i : 16 threads, thread (r, c) owns one element of the running 4x4 product
__global__ void pose_chain_fwd_kernel(const float* __restrict__ rel, int F, float* __restrict__ w2c) {
  __shared__ float cur[16];
  const int t = threadIdx.x, r = t >> 2, c = t & 3;
  cur[t] = (r == c) ? 1.0f : 0.0f;
  w2c[t] = cur[t];
  __syncthreads();
  for (int i = 0; i < F; ++i) {
    const float* A = rel + (int64_t)i * 16;
    const float v = A[r * 4] * cur[c] + A[r * 4 + 1] * cur[4 + c] + A[r * 4 + 2] * cur[8 + c] + A[r * 4 + 3] * cur[12 + c];
    __syncthreads();
    cur[t] = v;
    w2c[(int64_t)(i + 1) * 16 + t] = v;
    __syncthreads();
  }
}
// d_rel_i = G_{i+1} w2c_i^T,  G_i = d_w2c_i + rel_i^T G_{i+1}   (G: running adjoint of w2c_i)
__global__ void pose_chain_bwd_kernel(const float* __restrict__ rel, const float* __restrict__ w2c, int F,
                                      const float* __restrict__ d_w2c, float* __restrict__ d_rel) {
  __shared__ float G[16];
  const int t = threadIdx.x, r = t >> 2, c = t & 3;
  G[t] = d_w2c[(int64_t)F * 16 + t];
  __syncthreads();
  for (int i = F - 1; i >= 0; --i) {
    const float* A = rel + (int64_t)i * 16;
    const float* W = w2c + (int64_t)i * 16;
    d_rel[(int64_t)i * 16 + t] = G[r * 4] * W[c * 4] + G[r * 4 + 1] * W[c * 4 + 1] + G[r * 4 + 2] * W[c * 4 + 2] + G[r * 4 + 3] * W[c * 4 + 3];
    const float v = d_w2c[(int64_t)i * 16 + t] + A[r] * G[c] + A[4 + r] * G[4 + c] + A[8 + r] * G[8 + c] + A[12 + r] * G[12 + c];
    __syncthreads();
    G[t] = v;
    __syncthreads();
  }
}

}  // namespace
}  // namespace cope

using namespace cope;
extern "C" {

int cope_pose_integrate_fwd(const float* wv, const float* dt, int F, int n_sub, float* rel, cope_stream_t s) {
  if (F <= 0) return 0;
  COPE_REQUIRE(n_sub >= 1 && n_sub <= kMaxSub, "pose_integrate: n_sub=%d out of range [1,%d]", n_sub, kMaxSub);
  pose_integrate_fwd_kernel<<<(F + 63) / 64, 64, 0, as_stream(s)>>>(wv, dt, F, n_sub, rel);
  COPE_CHECK_LAUNCH("pose_integrate_fwd");
  return 0;
}

int cope_pose_integrate_bwd(const float* wv, const float* dt, int F, int n_sub, const float* d_rel, float* d_wv, float* d_dt,
                            cope_stream_t s) {
  if (F <= 0) return 0;
  COPE_REQUIRE(n_sub >= 1 && n_sub <= kMaxSub, "pose_integrate: n_sub=%d out of range [1,%d]", n_sub, kMaxSub);
  pose_integrate_bwd_kernel<<<(F + 31) / 32, 32, 0, as_stream(s)>>>(wv, dt, F, n_sub, d_rel, d_wv, d_dt);
  COPE_CHECK_LAUNCH("pose_integrate_bwd");
  return 0;
}

int cope_pose_chain_fwd(const float* rel, int F, float* w2c, cope_stream_t s) {
  COPE_REQUIRE(F >= 0, "pose_chain: F=%d", F);
  pose_chain_fwd_kernel<<<1, 16, 0, as_stream(s)>>>(rel, F, w2c);
  COPE_CHECK_LAUNCH("pose_chain_fwd");
  return 0;
}

int cope_pose_chain_bwd(const float* rel, const float* w2c, int F, const float* d_w2c, float* d_rel, cope_stream_t s) {
  if (F <= 0) return 0;
  pose_chain_bwd_kernel<<<1, 16, 0, as_stream(s)>>>(rel, w2c, F, d_w2c, d_rel);
  COPE_CHECK_LAUNCH("pose_chain_bwd");
  return 0;
}
}
