// Adam over ONE flat fp32 buffer (the optimiser step of train.py:59-60 / :532, torch.optim.Adam semantics, no amsgrad).
// torch's fused multi-tensor Adam walks ~80 parameter tensors in 64 K-element chunks: ~30 CTAs for the 1 M parameters of the
// SDF + colour + variance networks, two launches, ~50 us each in the ncu launch list of bench.py — 3.6 % of the training step for
// 28 MB of traffic.  With parameters, gradients and both moments flat (dist.FlatGradBucket.flatten_params_) the step is one
// elementwise grid over n elements.
//
//   m = m + (1 - b1) (g - m);  v = b2 v + (1 - b2) g g;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//
// `step` lives on the device (already incremented by the caller), so a CUDA-graph replay advances it like capturable Adam.
#include "common.cuh"

namespace cope {
namespace {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ step, float lr,
                                                        float beta1, float beta2, float eps, float weight_decay) {
  __shared__ float s_coef[2];
  if (threadIdx.x == 0) {
    const double t = (double)step[0];
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_coef[0] = (float)((double)lr / bc1);
    s_coef[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_coef[0], bc2_sqrt = s_coef[1];
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (weight_decay != 0.0f) gi = fmaf(weight_decay, pi, gi);        // L2 penalty folded into the gradient (torch.optim.Adam)
    const float mi = fmaf(omb1, gi - m[i], m[i]);                     // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(omb2 * gi, gi, beta2 * v[i]);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

}  // namespace
}  // namespace cope

using namespace cope;

extern "C" int cope_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* step,
                              float lr, float beta1, float beta2, float eps, float weight_decay, cope_stream_t s) {
  COPE_REQUIRE(n >= 0, "adam_step: n=%lld", (long long)n);
  if (n == 0) return 0;
  COPE_REQUIRE(param && grad && exp_avg && exp_avg_sq && step, "adam_step: null buffer");
  COPE_REQUIRE(beta1 >= 0.0f && beta1 < 1.0f && beta2 >= 0.0f && beta2 < 1.0f && eps >= 0.0f, "adam_step: betas (%g, %g) eps %g", beta1,
               beta2, eps);
  const int64_t blocks = ceil_div(n, (int64_t)256);
  adam_step_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(s)>>>(param, grad, exp_avg, exp_avg_sq, n, step,
                                                                                              lr, beta1, beta2, eps, weight_decay);
  COPE_CHECK_LAUNCH("adam_step");
  return 0;
}
