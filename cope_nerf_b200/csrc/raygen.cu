// Weight-norm prep, per-camera pose (so(3) exp + translation) and fused ray generation, forward and backward.
//   poses_retriever.py:25-32, common.py:255-308 (Exp / make_c2w), common.py:175-215 + training.py:474-487 (rays)
#include "common.cuh"

namespace cope {

// ------------------------------------------------------------------------------------ weight norm
// W = v * (g / ||v||_row)     one warp per output row
__global__ void weightnorm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ W,
                                      int rows, int cols) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const float* vr = v + (int64_t)row * cols;
  float ss = 0.0f;
  for (int c = lane; c < cols; c += 32) ss += vr[c] * vr[c];
  ss = warp_sum(ss);
  float sc = g[row] / sqrtf(ss);
  for (int c = lane; c < cols; c += 32) W[(int64_t)row * cols + c] = vr[c] * sc;
}

// dg = (dW . v)/||v|| ; dv = g/||v|| * dW - g (dW . v)/||v||^3 * v
__global__ void weightnorm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                      const float* __restrict__ dW, float* __restrict__ dv, float* __restrict__ dg,
                                      int rows, int cols) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  const float* vr = v + (int64_t)row * cols;
  const float* dr = dW + (int64_t)row * cols;
  float ss = 0.0f, dot = 0.0f;
  for (int c = lane; c < cols; c += 32) { ss += vr[c] * vr[c]; dot += vr[c] * dr[c]; }
  ss = warp_sum(ss); dot = warp_sum(dot);
  float nrm = sqrtf(ss), gi = g[row];
  float a = gi / nrm, b = gi * dot / (nrm * ss);
  for (int c = lane; c < cols; c += 32) dv[(int64_t)row * cols + c] = a * dr[c] - b * vr[c];
  if (lane == 0) dg[row] = dot / nrm;
}

// all layers of one network in ONE launch: flat = [W_0 | b_0 | W_1 | b_1 ...]
constexpr int kMaxWn = 16;
struct WnBatch {
  const float* v[kMaxWn]; const float* g[kMaxWn]; const float* b[kMaxWn];
  float* dv[kMaxWn]; float* dg[kMaxWn]; float* db[kMaxWn];
  int rows[kMaxWn], cols[kMaxWn], row0[kMaxWn];
  long long w_off[kMaxWn], b_off[kMaxWn];
  int n, total_rows;
};
__global__ void flat_weights_fwd_kernel(const __grid_constant__ WnBatch B, float* __restrict__ flat) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B.total_rows) return;
  int l = 0;
  while (l + 1 < B.n && row >= B.row0[l + 1]) ++l;
  row -= B.row0[l];
  const int lane = threadIdx.x & 31, cols = B.cols[l];
  const float* vr = B.v[l] + (int64_t)row * cols;
  float ss = 0.0f;
  for (int c = lane; c < cols; c += 32) ss += vr[c] * vr[c];
  ss = warp_sum(ss);
  const float sc = B.g[l][row] / sqrtf(ss);
  float* W = flat + B.w_off[l] + (int64_t)row * cols;
  for (int c = lane; c < cols; c += 32) W[c] = vr[c] * sc;
  if (lane == 0) flat[B.b_off[l] + row] = B.b[l][row];
}
__global__ void flat_weights_bwd_kernel(const __grid_constant__ WnBatch B, const float* __restrict__ dflat, int accumulate) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B.total_rows) return;
  int l = 0;
  while (l + 1 < B.n && row >= B.row0[l + 1]) ++l;
  row -= B.row0[l];
  const int lane = threadIdx.x & 31, cols = B.cols[l];
  const float* vr = B.v[l] + (int64_t)row * cols;
  const float* dr = dflat + B.w_off[l] + (int64_t)row * cols;
  float ss = 0.0f, dot = 0.0f;
  for (int c = lane; c < cols; c += 32) { ss += vr[c] * vr[c]; dot += vr[c] * dr[c]; }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const float nrm = sqrtf(ss), gi = B.g[l][row];
  const float a = gi / nrm, b = gi * dot / (nrm * ss);
  float* dv = B.dv[l] + (int64_t)row * cols;
  // accumulate: dv / dg / db ARE the parameters' .grad buffers (views of one flat bucket): add in place instead of handing
  // ~80 temporaries to autograd, which would launch one elementwise add per parameter tensor
  if (accumulate) {
    for (int c = lane; c < cols; c += 32) dv[c] += a * dr[c] - b * vr[c];
    if (lane == 0) { B.dg[l][row] += dot / nrm; B.db[l][row] += dflat[B.b_off[l] + row]; }
  } else {
    for (int c = lane; c < cols; c += 32) dv[c] = a * dr[c] - b * vr[c];
    if (lane == 0) { B.dg[l][row] = dot / nrm; B.db[l][row] = dflat[B.b_off[l] + row]; }
  }
}

// ------------------------------------------------------------------------------------ 4x4 helpers (fp64)
__device__ void inv4(const double* m, double* o) {
  double a[16];
  a[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  a[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  a[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  a[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  a[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  a[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  a[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  a[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  a[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  a[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  a[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  a[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  a[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  a[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  a[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  a[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * a[0] + m[1] * a[4] + m[2] * a[8] + m[3] * a[12];
  double id = 1.0 / det;
  for (int i = 0; i < 16; ++i) o[i] = a[i] * id;
}
__device__ void mul4(const double* a, const double* b, double* o) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      o[i * 4 + j] = s;
    }
}
__device__ void mul4_tn(const double* a, const double* b, double* o) {  // a^T b
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[k * 4 + i] * b[k * 4 + j];
      o[i * 4 + j] = s;
    }
}
__device__ void mul4_nt(const double* a, const double* b, double* o) {  // a b^T
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[j * 4 + k];
      o[i * 4 + j] = s;
    }
}
template <typename T>
__device__ void load16(const float* p, T* o) { for (int i = 0; i < 16; ++i) o[i] = (T)p[i]; }

// M = inv(scale) inv(world) inv(camera);   also returns the three inverses when asked
__device__ void unproject_matrix(const float* cam, const float* world, const float* scale, double* M, double* Wi,
                                 double* Si, double* Ki) {
  double c[16], w[16], s[16], ki[16], wi[16], si[16], t[16];
  load16(cam, c); load16(world, w); load16(scale, s);
  inv4(c, ki); inv4(w, wi); inv4(s, si);
  mul4(si, wi, t);
  mul4(t, ki, M);
  if (Wi) for (int i = 0; i < 16; ++i) { Wi[i] = wi[i]; Si[i] = si[i]; Ki[i] = ki[i]; }
}

// ------------------------------------------------------------------------------------ pose
// c2w = [Exp(r) | t; 0 0 0 1] @ init      (fp32, same formula order as common.py:268-288)
__global__ void pose_fwd_kernel(const float* __restrict__ r, const float* __restrict__ t, const float* __restrict__ init,
                                float* __restrict__ c2w) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float r0 = r[0], r1 = r[1], r2 = r[2];
  float K[9] = {0, -r2, r1, r2, 0, -r0, -r1, r0, 0};
  float n = sqrtf(r0 * r0 + r1 * r1 + r2 * r2) + 1e-15f;
  float A = sinf(n) / n, B = (1.0f - cosf(n)) / (n * n);
  float m[16];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float k2 = 0;
      for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
      m[i * 4 + j] = (i == j ? 1.0f : 0.0f) + A * K[i * 3 + j] + B * k2;
    }
  m[3] = t[0]; m[7] = t[1]; m[11] = t[2];
  m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = 0;
      for (int k = 0; k < 4; ++k) s += m[i * 4 + k] * init[k * 4 + j];
      c2w[i * 4 + j] = s;
    }
}

// Mirrors the autograd chain of Exp (incl. torch's zero sub-gradient of ||r|| at r = 0); fp64 internally.
__global__ void pose_bwd_kernel(const float* __restrict__ r, const float* __restrict__ t, const float* __restrict__ init,
                                const float* __restrict__ d_c2w, float* __restrict__ dr, float* __restrict__ dt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double dc[16], in[16], dM[16];
  load16(d_c2w, dc); load16(init, in);
  mul4_nt(dc, in, dM);                        // d(make_c2w) = d_c2w @ init^T
  dt[0] = (float)dM[3]; dt[1] = (float)dM[7]; dt[2] = (float)dM[11];
  double rv[3] = {r[0], r[1], r[2]};
  double K[9] = {0, -rv[2], rv[1], rv[2], 0, -rv[0], -rv[1], rv[0], 0};
  double K2[9], dR[9], dK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j];
      K2[i * 3 + j] = s;
      dR[i * 3 + j] = dM[i * 4 + j];
    }
  double nr = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  double n = nr + 1e-15;
  double A = sin(n) / n, B = (1.0 - cos(n)) / (n * n);
  double dA = 0, dB = 0;
  for (int i = 0; i < 9; ++i) { dA += dR[i] * K[i]; dB += dR[i] * K2[i]; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = A * dR[i * 3 + j];
      for (int k = 0; k < 3; ++k) s += B * (dR[i * 3 + k] * K[j * 3 + k] + K[k * 3 + i] * dR[k * 3 + j]);
      dK[i * 3 + j] = s;
    }
  double g[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
  if (nr > 0.0) {
    double dAn = cos(n) / n - sin(n) / (n * n);
    double dBn = sin(n) / (n * n) - 2.0 * (1.0 - cos(n)) / (n * n * n);
    double dn = dA * dAn + dB * dBn;
    for (int i = 0; i < 3; ++i) g[i] += dn * rv[i] / nr;
  }
  dr[0] = (float)g[0]; dr[1] = (float)g[1]; dr[2] = (float)g[2];
  (void)t;
}

// ------------------------------------------------------------------------------------ ray generation
__global__ void raygen_fwd_kernel(const float* __restrict__ pix, const float* __restrict__ cam,
                                  const float* __restrict__ world, const float* __restrict__ scale, int64_t N,
                                  float* __restrict__ ro, float* __restrict__ rd, float* __restrict__ rn) {
  __shared__ float sM[12];
  if (threadIdx.x == 0) {
    double M[16];
    unproject_matrix(cam, world, scale, M, nullptr, nullptr, nullptr);
    for (int i = 0; i < 12; ++i) sM[i] = (float)M[i];
  }
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x = pix[i * 2], y = pix[i * 2 + 1];
  float v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = sM[k * 4] * x + sM[k * 4 + 1] * y + sM[k * 4 + 2];
  float nv = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) { ro[i * 3 + k] = sM[k * 4 + 3]; rd[i * 3 + k] = v[k] / nv; }
  rn[i] = nv;
}

// per-ray partials of dL/dM (3x4) -> block reduce -> atomics into acc[12]
__global__ void raygen_bwd_kernel(const float* __restrict__ pix, const float* __restrict__ cam,
                                  const float* __restrict__ world, const float* __restrict__ scale, int64_t N,
                                  const float* __restrict__ d_ro, const float* __restrict__ d_rd,
                                  const float* __restrict__ d_rn, float* __restrict__ acc) {
  __shared__ float sM[12];
  __shared__ float red[12][8];
  if (threadIdx.x == 0) {
    double M[16];
    unproject_matrix(cam, world, scale, M, nullptr, nullptr, nullptr);
    for (int i = 0; i < 12; ++i) sM[i] = (float)M[i];
  }
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float part[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) part[k] = 0.0f;
  if (i < N) {
    float x = pix[i * 2], y = pix[i * 2 + 1];
    float v[3], d[3], gd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = sM[k * 4] * x + sM[k * 4 + 1] * y + sM[k * 4 + 2];
    float nv = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { d[k] = v[k] / nv; gd[k] = d_rd ? d_rd[i * 3 + k] : 0.0f; dot += d[k] * gd[k]; }
    float gn = d_rn ? d_rn[i] : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float dv = (gd[k] - d[k] * dot) / nv + gn * d[k];
      part[k * 4 + 0] = dv * x; part[k * 4 + 1] = dv * y; part[k * 4 + 2] = dv;
      part[k * 4 + 3] = d_ro ? d_ro[i * 3 + k] : 0.0f;
    }
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    float s = warp_sum(part[k]);
    if (lane == 0) red[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.0f;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += red[threadIdx.x][q];
    atomicAdd(acc + threadIdx.x, s);
  }
}

// dM (3x4 in acc) -> d world_mat:  M = Si Wi Ki, Wi = inv(W):  dWi = Si^T dM Ki^T ;  dW = -Wi^T dWi Wi^T
__global__ void raygen_bwd_finish_kernel(const float* __restrict__ cam, const float* __restrict__ world,
                                         const float* __restrict__ scale, const float* __restrict__ acc,
                                         float* __restrict__ d_world) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double M[16], Wi[16], Si[16], Ki[16], dM[16], t1[16], dWi[16], t2[16], dW[16];
  unproject_matrix(cam, world, scale, M, Wi, Si, Ki);
  for (int i = 0; i < 12; ++i) dM[i] = acc[i];
  for (int i = 12; i < 16; ++i) dM[i] = 0.0;
  mul4_tn(Si, dM, t1);
  mul4_nt(t1, Ki, dWi);
  mul4_tn(Wi, dWi, t2);
  mul4_nt(t2, Wi, dW);
  for (int i = 0; i < 16; ++i) d_world[i] = (float)(-dW[i]);
}

}  // namespace cope

using namespace cope;

extern "C" {

int cope_weightnorm_fwd(const float* v, const float* g, float* W, int rows, int cols, cope_stream_t s) {
  if (rows <= 0) return 0;
  weightnorm_fwd_kernel<<<(rows + 3) / 4, 128, 0, as_stream(s)>>>(v, g, W, rows, cols);
  COPE_CHECK_LAUNCH("weightnorm_fwd");
  return 0;
}
int cope_weightnorm_bwd(const float* v, const float* g, const float* dW, float* dv, float* dg, int rows, int cols,
                        cope_stream_t s) {
  if (rows <= 0) return 0;
  weightnorm_bwd_kernel<<<(rows + 3) / 4, 128, 0, as_stream(s)>>>(v, g, dW, dv, dg, rows, cols);
  COPE_CHECK_LAUNCH("weightnorm_bwd");
  return 0;
}

static int fill_wn(WnBatch& B, int n, const void* const* v, const void* const* g, const void* const* b, const int* rows,
                   const int* cols, const int64_t* w_off, const int64_t* b_off) {
  COPE_REQUIRE(n >= 1 && n <= kMaxWn, "flat_weights: %d layers (max %d)", n, kMaxWn);
  B.n = n; B.total_rows = 0;
  for (int l = 0; l < n; ++l) {
    B.v[l] = (const float*)v[l]; B.g[l] = (const float*)g[l]; B.b[l] = (const float*)b[l];
    B.rows[l] = rows[l]; B.cols[l] = cols[l]; B.row0[l] = B.total_rows; B.total_rows += rows[l];
    B.w_off[l] = w_off[l]; B.b_off[l] = b_off[l];
  }
  return 0;
}
int cope_flat_weights_fwd(int n, const void* const* v, const void* const* g, const void* const* b, const int* rows,
                          const int* cols, const int64_t* w_off, const int64_t* b_off, float* flat, cope_stream_t s) {
  WnBatch B{};
  if (int rc = fill_wn(B, n, v, g, b, rows, cols, w_off, b_off)) return rc;
  flat_weights_fwd_kernel<<<(B.total_rows + 3) / 4, 128, 0, as_stream(s)>>>(B, flat);
  COPE_CHECK_LAUNCH("flat_weights_fwd");
  return 0;
}
int cope_flat_weights_bwd(int n, const void* const* v, const void* const* g, const int* rows, const int* cols,
                          const int64_t* w_off, const int64_t* b_off, const float* dflat, void* const* dv, void* const* dg,
                          void* const* db, int accumulate, cope_stream_t s) {
  WnBatch B{};
  if (int rc = fill_wn(B, n, v, g, g, rows, cols, w_off, b_off)) return rc;
  for (int l = 0; l < n; ++l) { B.dv[l] = (float*)dv[l]; B.dg[l] = (float*)dg[l]; B.db[l] = (float*)db[l]; }
  flat_weights_bwd_kernel<<<(B.total_rows + 3) / 4, 128, 0, as_stream(s)>>>(B, dflat, accumulate);
  COPE_CHECK_LAUNCH("flat_weights_bwd");
  return 0;
}

int cope_pose_fwd(const float* r, const float* t, const float* init_c2w, float* c2w, cope_stream_t s) {
  pose_fwd_kernel<<<1, 32, 0, as_stream(s)>>>(r, t, init_c2w, c2w);
  COPE_CHECK_LAUNCH("pose_fwd");
  return 0;
}
int cope_pose_bwd(const float* r, const float* t, const float* init_c2w, const float* d_c2w, float* dr, float* dt,
                  cope_stream_t s) {
  pose_bwd_kernel<<<1, 32, 0, as_stream(s)>>>(r, t, init_c2w, d_c2w, dr, dt);
  COPE_CHECK_LAUNCH("pose_bwd");
  return 0;
}

int cope_raygen_fwd(const float* pixels, const float* camera_mat, const float* world_mat, const float* scale_mat,
                    int64_t N, float* rays_o, float* rays_d, float* rays_d_norm, cope_stream_t s) {
  if (N <= 0) return 0;
  raygen_fwd_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(s)>>>(pixels, camera_mat, world_mat, scale_mat, N,
                                                                         rays_o, rays_d, rays_d_norm);
  COPE_CHECK_LAUNCH("raygen_fwd");
  return 0;
}

int cope_raygen_bwd(const float* pixels, const float* camera_mat, const float* world_mat, const float* scale_mat,
                    int64_t N, const float* d_rays_o, const float* d_rays_d, const float* d_norm, float* d_world,
                    float* ws, cope_stream_t s_) {
  cudaStream_t s = as_stream(s_);
  cudaMemsetAsync(ws, 0, 16 * sizeof(float), s);
  if (N > 0) {
    raygen_bwd_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, s>>>(pixels, camera_mat, world_mat, scale_mat, N, d_rays_o,
                                                                d_rays_d, d_norm, ws);
    COPE_CHECK_LAUNCH("raygen_bwd");
  }
  raygen_bwd_finish_kernel<<<1, 32, 0, s>>>(camera_mat, world_mat, scale_mat, ws, d_world);
  COPE_CHECK_LAUNCH("raygen_bwd_finish");
  return 0;
}

}  // extern "C"
