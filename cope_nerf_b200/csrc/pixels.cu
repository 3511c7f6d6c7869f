// Training-pixel selection of process_data (model/training.py:413-471) as ONE launch: patch corners -> flat pixel ids ->
// integer and normalised pixel coordinates -> target colours gathered from the frame.  The reference draws the corners with a
// CPU randperm over (h-ps+1)(w-ps+1) elements and builds the full h*w pixel grid (arange_pixels, model/common.py:12-39)
// every step.  Corners either come from the caller (the reference's own randperm stream: bit-exact ids) or, with
// corners == NULL, from a keyed bijection of [0, M) evaluated on the device (distinct corners, no sort, no host work).
#include "common.cuh"

namespace cope {
namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// keyed permutation of [0, M): 4-round balanced Feistel network on 2*hb bits (2^(2hb) >= M) with cycle walking, so
// perm(0), perm(1), ... are distinct elements of [0, M) -- sampling without replacement in O(1) per draw
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t M, int hb, uint64_t seed) {
  const uint32_t mask = (1u << hb) - 1u;
  uint64_t x = i;
  do {
    uint32_t l = (uint32_t)(x >> hb) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t f = mix32(r ^ (uint32_t)(seed >> (16 * (k & 1))) ^ (0x9e3779b9U * (k + 1))) & mask;
      const uint32_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = ((uint64_t)l << hb) | r;
  } while (x >= M);
  return x;
}

__global__ void __launch_bounds__(256) sample_pixels_kernel(const int64_t* __restrict__ corners, uint64_t seed, int hb, int h, int w,
                                                            int ps, int n_patches, const float* __restrict__ img,
                                                            int64_t* __restrict__ ray_idx, float* __restrict__ pix,
                                                            float* __restrict__ npix, float* __restrict__ rgb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int pp = ps * ps;
  if (i >= n_patches * pp) return;
  const int k = i / pp, o = i - k * pp;
  const int wa = w - ps + 1;
  const int64_t M = (int64_t)(h - ps + 1) * wa;
  const int64_t c = corners ? corners[k] : (int64_t)feistel_perm((uint64_t)k, (uint64_t)M, hb, seed);
  const int row = (int)(c / wa) + o / ps, col = (int)(c % wa) + o % ps;       // row-major inside the patch (:428-435)
  const int64_t id = (int64_t)row * w + col;
  if (ray_idx) ray_idx[i] = id;
  if (pix) { pix[2 * i] = (float)col; pix[2 * i + 1] = (float)row; }
  if (npix) {                                                                  // model/common.py:35-38, same rounding
    npix[2 * i] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)col), (float)(w - 1)), 1.0f);
    npix[2 * i + 1] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)row), (float)(h - 1)), 1.0f);
  }
  if (rgb && img) {
    const int64_t hw = (int64_t)h * w;
    rgb[3 * i] = img[id]; rgb[3 * i + 1] = img[hw + id]; rgb[3 * i + 2] = img[2 * hw + id];
  }
}

}  // namespace
}  // namespace cope

using namespace cope;

extern "C" int cope_sample_pixels(const int64_t* corners, uint64_t seed, int h, int w, int patch_size, int n_patches,
                                  const float* img, int64_t* ray_idx, float* pix, float* norm_pix, float* rgb_gt, cope_stream_t s) {
  COPE_REQUIRE(patch_size >= 1 && h >= patch_size && w >= patch_size && h >= 2 && w >= 2, "sample_pixels: h=%d w=%d patch=%d", h, w,
               patch_size);
  const int64_t M = (int64_t)(h - patch_size + 1) * (w - patch_size + 1);
  COPE_REQUIRE(n_patches >= 0 && n_patches <= M, "sample_pixels: %d patches out of %lld corners", n_patches, (long long)M);
  if (n_patches == 0) return 0;
  int hb = 1;
  while ((1ll << (2 * hb)) < M) ++hb;
  const int n = n_patches * patch_size * patch_size;
  sample_pixels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(s)>>>(corners, seed, hb, h, w, patch_size, n_patches, img, ray_idx,
                                                                             pix, norm_pix, rgb_gt);
  COPE_CHECK_LAUNCH("sample_pixels");
  return 0;
}
