// Phase timeline of k_shade_bwd_fused_ws (CTA 0, tiles 4..7 of its range): where do the ~29 k cycles per 128-sample tile go?
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DUBN_WS_TRACE -I include scripts/trace_ws.cu -o scripts/_bin/trace_ws
// Prints, per tile and warp, clock64 deltas relative to the row-warp-0 tile start.  Not part of the library or the tests.
#include "../unboundednerfpytorch_b200/csrc/shade_tc.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace ubn { thread_local cudaError_t g_last_error = cudaSuccess; void count_launch() {} }

int main(int argc, char** argv) {
  const int flags = argc > 1 ? atoi(argv[1]) : 0;     // single_pass bits of ubn_rgbnet_bwd_tc_fused (4 = panel-layout saves)
  const int64_t tiles_per_cta = 12, n = 148 * tiles_per_cta * 128, n_rays = n / 512;
  auto dev = [](size_t bytes) { void* p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes); return p; };
  std::vector<float> h(n * 128);
  srand(1);
  auto fill = [&](float* d, size_t cnt, float lo, float hi) {
    for (size_t i = 0; i < cnt; ++i) h[i] = lo + (hi - lo) * (rand() / (float)RAND_MAX);
    cudaMemcpy(d, h.data(), cnt * 4, cudaMemcpyHostToDevice);
  };
  float *feat = (float*)dev(n * 12 * 4), *W1k = (float*)dev(128 * 12 * 4), *W2 = (float*)dev(128 * 128 * 4), *W3 = (float*)dev(3 * 128 * 4);
  float *rgb = (float*)dev(n * 3 * 4), *h1 = (float*)dev(n * 128 * 4), *h2 = (float*)dev(n * 128 * 4), *g = (float*)dev(n * 3 * 4);
  float *gfeat = (float*)dev(n * 12 * 4), *gvb = (float*)dev(n_rays * 128 * 4), *gW1 = (float*)dev(128 * 12 * 4), *gW2 = (float*)dev(128 * 128 * 4);
  float *gb2 = (float*)dev(128 * 4), *gW3 = (float*)dev(3 * 128 * 4), *gb3 = (float*)dev(16);
  int64_t* ray = (int64_t*)dev(n * 8);
  uint32_t* masks = (uint32_t*)dev((n / 128) * 512 * 4);
  uint32_t* masks1 = (uint32_t*)dev((n / 128) * 512 * 4);
  cudaMemset(masks1, 0x5a, (n / 128) * 512 * 4);
  std::vector<int64_t> hr(n);
  for (int64_t i = 0; i < n; ++i) hr[i] = i / 512;
  cudaMemcpy(ray, hr.data(), n * 8, cudaMemcpyHostToDevice);
  fill(feat, n * 12, -1, 1); fill(W1k, 128 * 12, -.3f, .3f); fill(W2, 128 * 128, -.1f, .1f); fill(W3, 3 * 128, -.1f, .1f);
  fill(rgb, n * 3, .05f, .95f); fill(h1, n * 128, -1, 1); fill(h2, n * 128, -1, 1); fill(g, n * 3, -1, 1);
  for (int rep = 0; rep < 3; ++rep) {
    int rc = ubn_rgbnet_bwd_tc_fused(feat, ray, W1k, W2, W3, rgb, h1, h2, g, n, gfeat, gvb, gW1, gW2, gb2, gW3, gb3, masks, masks1, flags, nullptr);
    if (rc) { printf("launch rc=%d\n", rc); return 1; }
    cudaDeviceSynchronize();
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  ubn_rgbnet_bwd_tc_fused(feat, ray, W1k, W2, W3, rgb, h1, h2, g, n, gfeat, gvb, gW1, gW2, gb2, gW3, gb3, masks, masks1, flags, nullptr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("both kernels: %.3f ms for %lld tiles/CTA -> %.2f us per tile (incl. dW2 kernel)\n", ms, (long long)tiles_per_cta, ms * 1e3 / tiles_per_cta);
  static long long t[4 * 8 * 64];
  cudaMemcpyFromSymbol(t, tc::g_ws_trace, sizeof(t));
  const char* rown[17] = {"start", "inputs", "A0", "A1", "A2", "A3", "syncA", "mma1_issued", "mma1_done", "B0", "B1", "B2", "B3", "syncB", "mma2_issued", "mma2_done", "end"};
  for (int tt = 0; tt < 4; ++tt) {
    const long long t0 = t[(tt * 8 + 0) * 64 + 0];
    printf("== tile %d (t0 = row warp 0 start; previous tile end at %lld)\n", tt + 4, tt ? t[((tt - 1) * 8 + 0) * 64 + 16] - t0 : 0LL);
    for (int w = 0; w < 4; w += 3) {
      printf(" row warp %d:", w);
      for (int s = 0; s < 17; ++s) printf(" %s=%lld", rown[s], t[(tt * 8 + w) * 64 + s] - t0);
      printf("\n");
    }
    for (int w = 4; w < 8; w += 3) {
      printf(" col warp %d: ", w);
      for (int c = 0; c < 8; ++c)
        printf("[c%d wait@%lld got@%lld rel@%lld] ", c, t[(tt * 8 + w) * 64 + c] - t0, t[(tt * 8 + w) * 64 + 8 + c] - t0, t[(tt * 8 + w) * 64 + 16 + c] - t0);
      printf("\n");
    }
  }
  return 0;
}
