// grid_sweep.cu -- whole-grid streaming sweeps of libubnerf_b200.so: total-variation gradient, the three
// Adam variants, and the fused TV + Adam + grad-zeroing training-step tail.
//
// Replaces total_variation_kernel.cu:14-67 (K20) and adam_upd_kernel.cu:9-132 (K17-K19) of the
// reference.  These sweeps move more bytes per training step than the ray work at 320^3 (SURVEY.md 8a
// a11/a12), so they are written as HBM-streaming kernels: 32-bit magic-number index decomposition
// (no 64-bit div/mod per element), 128-bit accesses where the layout allows, and no write traffic for
// elements masked Adam leaves untouched.
#include "common.cuh"

namespace ubn {

// ---- fast unsigned division by a runtime constant (n < 2^31, d >= 1) ---------------------------
struct FastDiv {
  uint32_t d, mul, shr;
};

static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d == 1) { f.mul = 0; f.shr = 0; return f; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;           // l = ceil(log2 d)
  const uint64_t m = (((1ull << 32) * ((1ull << l) - d)) / d) + 1;
  f.mul = (uint32_t)m;
  f.shr = l;
  return f;
}

__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  if (f.d == 1) return n;
  const uint32_t t = __umulhi(n, f.mul);
  return (t + ((n - t) >> 1)) >> (f.shr - 1);
}

struct GridShape {
  FastDiv inner, k, j, i;   // divisors: inner, sz_k, sz_j, sz_i
  int64_t n;                // total elements
  int64_t sk, sj, si;       // neighbour strides in elements
};

static GridShape make_shape(int64_t lead, int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t inner) {
  GridShape g;
  g.inner = make_fastdiv((uint32_t)inner);
  g.k = make_fastdiv((uint32_t)sz_k);
  g.j = make_fastdiv((uint32_t)sz_j);
  g.i = make_fastdiv((uint32_t)sz_i);
  g.n = lead * sz_i * sz_j * sz_k * inner;
  g.sk = inner;
  g.sj = sz_k * inner;
  g.si = sz_j * sz_k * inner;
  return g;
}

struct Ijk {
  uint32_t i, j, k;
};

template <bool kWide>
__device__ __forceinline__ Ijk decompose(int64_t m, const GridShape& g) {
  Ijk r;
  if (!kWide) {
    const uint32_t v = fdiv((uint32_t)m, g.inner);
    const uint32_t q1 = fdiv(v, g.k);
    r.k = v - q1 * g.k.d;
    const uint32_t q2 = fdiv(q1, g.j);
    r.j = q1 - q2 * g.j.d;
    const uint32_t q3 = fdiv(q2, g.i);
    r.i = q2 - q3 * g.i.d;
  } else {
    const int64_t v = m / g.inner.d;
    r.k = (uint32_t)(v % g.k.d);
    r.j = (uint32_t)(v / g.k.d % g.j.d);
    r.i = (uint32_t)(v / g.k.d / g.j.d % g.i.d);
  }
  return r;
}

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// TV gradient of one element, reference order and weights (k: wz, j: wy, i: wz -- sic, :27-32)
__device__ __forceinline__ float tv_term(const float* __restrict__ param, int64_t m, const Ijk& c, const GridShape& g,
                                         float wy, float wz) {
  const float p = param[m];
  float add = 0;
  add += (c.k == 0 ? 0 : wz * clamp1(p - param[m - g.sk]));
  add += (c.k == g.k.d - 1 ? 0 : wz * clamp1(p - param[m + g.sk]));
  add += (c.j == 0 ? 0 : wy * clamp1(p - param[m - g.sj]));
  add += (c.j == g.j.d - 1 ? 0 : wy * clamp1(p - param[m + g.sj]));
  add += (c.i == 0 ? 0 : wz * clamp1(p - param[m - g.si]));
  add += (c.i == g.i.d - 1 ? 0 : wz * clamp1(p - param[m + g.si]));
  return add;
}

template <bool kDense, bool kWide>
__global__ void __launch_bounds__(256) k_total_variation(const float* __restrict__ param, float* __restrict__ grad,
                                                         float wy, float wz, GridShape g) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.n) return;
  const float gr = grad[m];
  if (!(kDense || gr != 0)) return;
  const Ijk c = decompose<kWide>(m, g);
  grad[m] = gr + tv_term(param, m, c, g, wy, wz);
}

// ---- Adam (adam_upd_kernel.cu:14-16, 34-38, 52-56) ----------------------------------------------
struct AdamHyper {
  float step_size, beta1, beta2, eps;
};

static AdamHyper make_hyper(int step, float beta1, float beta2, float lr, float eps) {
  AdamHyper h;
  // host-side float arithmetic of the reference (adam_upd_kernel.cu:72)
  h.step_size = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
  h.beta1 = beta1;
  h.beta2 = beta2;
  h.eps = eps;
  return h;
}

template <int kMode>
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float perlr, const AdamHyper& h) {
  // Spelled with explicit roundings so that the fma contraction is the one nvcc picks for the reference's
  // expressions (verified in its SASS: m' = fma(m, b1, (1-b1)*g); v' = fma(v, b2, ((1-b2)*g)*g)) -> bit-identical
  // parameters / moments to the reference extension on the same inputs.
  m = __fmaf_rn(m, h.beta1, __fmul_rn(__fsub_rn(1.f, h.beta1), g));
  v = __fmaf_rn(v, h.beta2, __fmul_rn(__fmul_rn(__fsub_rn(1.f, h.beta2), g), g));
  const float den = __fadd_rn(__fsqrt_rn(v), h.eps);
  if (kMode == 2) p = __fsub_rn(p, __fdiv_rn(__fmul_rn(__fmul_rn(h.step_size, perlr), m), den));
  else            p = __fsub_rn(p, __fdiv_rn(__fmul_rn(h.step_size, m), den));
}

// vectorised (float4) main body + scalar tail
template <int kMode>
__global__ void __launch_bounds__(256) k_adam_vec4(float4* __restrict__ param, const float4* __restrict__ grad,
                                                   float4* __restrict__ exp_avg, float4* __restrict__ exp_avg_sq,
                                                   const float4* __restrict__ perlr, int64_t n4, AdamHyper h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = grad[i];
  if (kMode == 1 && g.x == 0 && g.y == 0 && g.z == 0 && g.w == 0) return;   // nothing to read or write
  float4 p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
  float4 lr4 = make_float4(0, 0, 0, 0);
  if (kMode == 2) lr4 = perlr[i];
  if (kMode != 1 || g.x != 0) adam_one<kMode>(p.x, g.x, m.x, v.x, lr4.x, h);
  if (kMode != 1 || g.y != 0) adam_one<kMode>(p.y, g.y, m.y, v.y, lr4.y, h);
  if (kMode != 1 || g.z != 0) adam_one<kMode>(p.z, g.z, m.z, v.z, lr4.z, h);
  if (kMode != 1 || g.w != 0) adam_one<kMode>(p.w, g.w, m.w, v.w, lr4.w, h);
  param[i] = p;
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
}

template <int kMode>
__global__ void __launch_bounds__(256) k_adam_scalar(float* __restrict__ param, const float* __restrict__ grad,
                                                     float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                     const float* __restrict__ perlr, int64_t begin, int64_t n,
                                                     AdamHyper h) {
  const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad[i];
  if (kMode == 1 && g == 0) return;
  float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
  adam_one<kMode>(p, g, m, v, kMode == 2 ? perlr[i] : 0.f, h);
  param[i] = p;
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
}

template <int kMode>
static int launch_adam(float* param, const float* grad, float* m, float* v, const float* perlr, int64_t n,
                       const AdamHyper& h, cudaStream_t st) {
  const bool aligned = ((((uintptr_t)param) | ((uintptr_t)grad) | ((uintptr_t)m) | ((uintptr_t)v) |
                         ((uintptr_t)perlr)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    k_adam_vec4<kMode><<<blocks_for(n4, 256), 256, 0, st>>>((float4*)param, (const float4*)grad, (float4*)m,
                                                           (float4*)v, (const float4*)perlr, n4, h);
    UBN_LAUNCH_CHECK();
  }
  const int64_t done = n4 * 4;
  if (done < n) {
    k_adam_scalar<kMode><<<blocks_for(n - done, 256), 256, 0, st>>>(param, grad, m, v, perlr, done, n, h);
    UBN_LAUNCH_CHECK();
  }
  return 0;
}

// ---- fused training-step tail: TV + Adam + grad zeroing -------------------------------------------
// The TV term of element m needs the OLD parameter values of its six neighbours while Adam overwrites
// parameters, so a literal single pass would race; without double-buffering the parameters the tail is
// two launches: (1) TV into grad (gated exactly like K20), (2) Adam (gated like K17/K18) that also
// writes grad <- 0 for the elements it consumed.  The win over the reference's three passes
// (fresh zero-filled grad allocation + TV + Adam) is that inactive elements (grad == 0, the vast
// majority late in training) cost one 4-byte read and nothing else, and no separate memset exists.
template <int kMode, bool kZero>
__global__ void __launch_bounds__(256) k_adam_zero_vec4(float4* __restrict__ param, float4* __restrict__ grad,
                                                        float4* __restrict__ exp_avg,
                                                        float4* __restrict__ exp_avg_sq, int64_t n4, AdamHyper h) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 g = grad[i];
  const bool any = (g.x != 0) | (g.y != 0) | (g.z != 0) | (g.w != 0);
  if (kMode == 1 && !any) return;
  float4 p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
  if (kMode != 1 || g.x != 0) adam_one<kMode>(p.x, g.x, m.x, v.x, 0.f, h);
  if (kMode != 1 || g.y != 0) adam_one<kMode>(p.y, g.y, m.y, v.y, 0.f, h);
  if (kMode != 1 || g.z != 0) adam_one<kMode>(p.z, g.z, m.z, v.z, 0.f, h);
  if (kMode != 1 || g.w != 0) adam_one<kMode>(p.w, g.w, m.w, v.w, 0.f, h);
  param[i] = p;
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
  if (kZero && any) grad[i] = make_float4(0, 0, 0, 0);
}

template <int kMode, bool kZero>
__global__ void __launch_bounds__(256) k_adam_zero_scalar(float* __restrict__ param, float* __restrict__ grad,
                                                          float* __restrict__ exp_avg,
                                                          float* __restrict__ exp_avg_sq, int64_t begin, int64_t n,
                                                          AdamHyper h) {
  const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = grad[i];
  if (kMode == 1 && g == 0) return;
  float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
  adam_one<kMode>(p, g, m, v, 0.f, h);
  param[i] = p;
  exp_avg[i] = m;
  exp_avg_sq[i] = v;
  if (kZero && g != 0) grad[i] = 0.f;
}

template <int kMode, bool kZero>
static int launch_adam_zero(float* param, float* grad, float* m, float* v, int64_t n, const AdamHyper& h,
                            cudaStream_t st) {
  const bool aligned = ((((uintptr_t)param) | ((uintptr_t)grad) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    k_adam_zero_vec4<kMode, kZero><<<blocks_for(n4, 256), 256, 0, st>>>((float4*)param, (float4*)grad, (float4*)m,
                                                                       (float4*)v, n4, h);
    UBN_LAUNCH_CHECK();
  }
  const int64_t done = n4 * 4;
  if (done < n) {
    k_adam_zero_scalar<kMode, kZero><<<blocks_for(n - done, 256), 256, 0, st>>>(param, grad, m, v, done, n, h);
    UBN_LAUNCH_CHECK();
  }
  return 0;
}

static int launch_tv(const float* param, float* grad, float wy, float wz, const GridShape& g, bool dense,
                     cudaStream_t st) {
  const bool wide = g.n >= (1ll << 31);
  const unsigned nb = blocks_for(g.n, 256);
  if (dense) {
    if (wide) k_total_variation<true, true><<<nb, 256, 0, st>>>(param, grad, wy, wz, g);
    else      k_total_variation<true, false><<<nb, 256, 0, st>>>(param, grad, wy, wz, g);
  } else {
    if (wide) k_total_variation<false, true><<<nb, 256, 0, st>>>(param, grad, wy, wz, g);
    else      k_total_variation<false, false><<<nb, 256, 0, st>>>(param, grad, wy, wz, g);
  }
  UBN_LAUNCH_CHECK();
  return 0;
}

}  // namespace ubn

using namespace ubn;

extern "C" {

int ubn_total_variation_add_grad(const float* param, float* grad, float wx, float wy, float wz, int64_t lead,
                                 int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t inner, int dense_mode,
                                 void* stream) {
  (void)wx;  // unused by the reference as well (total_variation_kernel.cu:31-32)
  const GridShape g = make_shape(lead, sz_i, sz_j, sz_k, inner);
  if (g.n <= 0) return 0;
  wy /= 6;   // host-side pre-division, total_variation_kernel.cu:45-47
  wz /= 6;
  return launch_tv(param, grad, wy, wz, g, dense_mode != 0, as_stream(stream));
}

int ubn_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* perlr, int64_t n,
                 int step, float beta1, float beta2, float lr, float eps, int mode, void* stream) {
  if (n <= 0) return 0;
  const AdamHyper h = make_hyper(step, beta1, beta2, lr, eps);
  cudaStream_t st = as_stream(stream);
  switch (mode) {
    case 0: return launch_adam<0>(param, grad, exp_avg, exp_avg_sq, nullptr, n, h, st);
    case 1: return launch_adam<1>(param, grad, exp_avg, exp_avg_sq, nullptr, n, h, st);
    case 2: return launch_adam<2>(param, grad, exp_avg, exp_avg_sq, perlr, n, h, st);
    default: return finish(cudaErrorInvalidValue);
  }
}

int ubn_tv_adam_fused(float* param, float* grad, float* exp_avg, float* exp_avg_sq, float wx, float wy, float wz,
                      int64_t lead, int64_t sz_i, int64_t sz_j, int64_t sz_k, int64_t inner, int tv_mode, int step,
                      float beta1, float beta2, float lr, float eps, int adam_mode, int zero_grad, void* stream) {
  (void)wx;
  const GridShape g = make_shape(lead, sz_i, sz_j, sz_k, inner);
  if (g.n <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (tv_mode != 0) {
    const int e = launch_tv(param, grad, wy / 6, wz / 6, g, tv_mode == 1, st);
    if (e) return e;
  }
  const AdamHyper h = make_hyper(step, beta1, beta2, lr, eps);
  if (adam_mode == 1)
    return zero_grad ? launch_adam_zero<1, true>(param, grad, exp_avg, exp_avg_sq, g.n, h, st)
                     : launch_adam_zero<1, false>(param, grad, exp_avg, exp_avg_sq, g.n, h, st);
  if (adam_mode == 0)
    return zero_grad ? launch_adam_zero<0, true>(param, grad, exp_avg, exp_avg_sq, g.n, h, st)
                     : launch_adam_zero<0, false>(param, grad, exp_avg, exp_avg_sq, g.n, h, st);
  return finish(cudaErrorInvalidValue);
}

}  // extern "C"
