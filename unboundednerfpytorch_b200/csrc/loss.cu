// loss.cu -- the always-on training losses of run_train.py:254-279 and their gradients in two launches
// ("in-kernel losses"):
//   loss = w_main * mse(rgb_marched, target) + w_entropy * entropy_last(alphainv_last) + w_rgbper * rgbper
//   mse      = mean_{r,c} (rgb_marched - target)^2                                     F.mse_loss, :254
//   entropy  = mean_r  -(p log p + (1-p) log(1-p)),  p = clamp(alphainv_last, 1e-6, 1-1e-6)   :258-261
//   rgbper   = sum_m weights_m * sum_c (raw_rgb_m - target[ray_id_m])^2 / n_rays        :275-278 (weights detached)
// The reference spends ~25 torch kernels on this (gather target[ray_id] -> 50 MB, sub, pow, sum, mul, sum, clamp, logs,
// means, and their autograd mirrors).  Here one grid-stride kernel reads every operand once, writes the three
// gradients and per-block partial sums (double), and a one-block kernel adds the partials in a fixed order, so the
// loss value is deterministic.
#include "common.cuh"

namespace ubn {

constexpr int kLossBlocks = 592;      // 4 per SM
constexpr int kLossThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLossThreads / 32; ++i) t += sh[i];
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(kLossThreads) k_render_loss(
    const float* __restrict__ rgb_marched, const float* __restrict__ alphainv_last, const float* __restrict__ raw_rgb,
    const float* __restrict__ weights, const int64_t* __restrict__ ray_id, const float* __restrict__ target, int64_t n_rays,
    int64_t n_pts, float w_main, float w_entropy, float w_rgbper, float* __restrict__ g_rgb_marched,
    float* __restrict__ g_alphainv, float* __restrict__ g_raw_rgb, double* __restrict__ partial) {
  __shared__ double sh[kLossThreads / 32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s_mse = 0, s_ent = 0, s_per = 0;
  const float inv_n = 1.f / (float)n_rays;
  const float g_mse = w_main * 2.f / (3.f * (float)n_rays);
  for (int64_t r = t0; r < n_rays; r += stride) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = rgb_marched[3 * r + c] - target[3 * r + c];
      s_mse += (double)(d * d);
      if (g_rgb_marched) g_rgb_marched[3 * r + c] = g_mse * d;
    }
    if (alphainv_last) {
      const float a = alphainv_last[r];
      const float p = fminf(fmaxf(a, 1e-6f), 1.f - 1e-6f);
      const float lp = logf(p), lq = logf(1.f - p);
      s_ent += (double)(-(p * lp + (1.f - p) * lq));
      // d/dp = -(log p - log(1-p)); clamp passes the gradient on [min, max] inclusive
      if (g_alphainv) g_alphainv[r] = (a >= 1e-6f && a <= 1.f - 1e-6f) ? w_entropy * inv_n * (lq - lp) : 0.f;
    }
  }
  if (raw_rgb) {
    const float g_per = w_rgbper * 2.f * inv_n;
    for (int64_t m = t0; m < n_pts; m += stride) {
      const int64_t r = ray_id[m];
      const float w = weights[m];
      float per = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = raw_rgb[3 * m + c] - target[3 * r + c];
        per = fmaf(d, d, per);
        if (g_raw_rgb) g_raw_rgb[3 * m + c] = g_per * w * d;
      }
      s_per += (double)(per * w);
    }
  }
  const double a = block_sum(s_mse, sh);
  const double b = block_sum(s_ent, sh);
  const double c = block_sum(s_per, sh);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = a;
    partial[3 * blockIdx.x + 1] = b;
    partial[3 * blockIdx.x + 2] = c;
  }
}

__global__ void k_render_loss_finish(const double* __restrict__ partial, int n_blocks, int64_t n_rays, float w_main,
                                     float w_entropy, float w_rgbper, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0, b = 0, c = 0;
  for (int i = 0; i < n_blocks; ++i) { a += partial[3 * i]; b += partial[3 * i + 1]; c += partial[3 * i + 2]; }
  const float mse = (float)(a / (3.0 * (double)n_rays));
  const float ent = (float)(b / (double)n_rays);
  const float per = (float)(c / (double)n_rays);
  out[0] = w_main * mse + w_entropy * ent + w_rgbper * per;
  out[1] = mse; out[2] = ent; out[3] = per;
}

}  // namespace ubn

using namespace ubn;

extern "C" int ubn_render_loss(const float* rgb_marched, const float* alphainv_last, const float* raw_rgb,
                               const float* weights, const int64_t* ray_id, const float* target, int64_t n_rays,
                               int64_t n_pts, float w_main, float w_entropy, float w_rgbper, float* out4,
                               float* grad_rgb_marched, float* grad_alphainv_last, float* grad_raw_rgb, double* scratch,
                               int64_t scratch_len, void* stream) {
  if (n_rays <= 0) return finish(cudaErrorInvalidValue);
  if (scratch_len < 3 * kLossBlocks) return finish(cudaErrorInvalidValue);
  cudaStream_t st = as_stream(stream);
  const bool ent = alphainv_last != nullptr && w_entropy != 0.f;
  const bool per = raw_rgb != nullptr && w_rgbper != 0.f && n_pts > 0;
  k_render_loss<<<kLossBlocks, kLossThreads, 0, st>>>(rgb_marched, ent ? alphainv_last : nullptr, per ? raw_rgb : nullptr, weights,
                                                     ray_id, target, n_rays, n_pts, w_main, w_entropy, w_rgbper,
                                                     grad_rgb_marched, ent ? grad_alphainv_last : nullptr,
                                                     per ? grad_raw_rgb : nullptr, scratch);
  UBN_LAUNCH_CHECK();
  k_render_loss_finish<<<1, 32, 0, st>>>(scratch, kLossBlocks, n_rays, w_main, ent ? w_entropy : 0.f, per ? w_rgbper : 0.f, out4);
  UBN_LAUNCH_CHECK();
  return 0;
}

