// probe_mma_rate.cu -- hardware probe (B200, sm_100a): issue rate of tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) as a
// function of the shared-memory operand layout, and a correctness check of the K-major SWIZZLE_128B layout.
//   layout 0 = K-major, no swizzle (8 x 16 B core matrices, LBO = 2048, SBO = 128: what shade_tc.cu used up to round 2)
//   layout 1 = K-major, SWIZZLE_128B (rows of 128 B = 32 tf32 of K, 8-row atoms of 1024 B, SBO = 1024, 16-B chunk ^= row & 7)
// Prints cycles per MMA for SS / TS forms, N = 128 / 256 / 16, and max |D - ref| for both layouts.
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a scripts/probe_mma_rate.cu -o scripts/_bin/probe_mma_rate
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kM = 128, kK = 32;
constexpr uint32_t kOpBytes = 32 * 1024;          // room for N = 256 rows x 128 B

__host__ __device__ inline uint32_t elem_offset(int layout, int r, int k) {
  if (layout == 0) return (uint32_t)(k >> 2) * 4096u + (uint32_t)r * 16u + (uint32_t)(k & 3) * 4u;     // panel stride 4096 (256 rows)
  uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)k * 4u;
  return off ^ (((off >> 7) & 7u) << 4);
}
__device__ inline uint64_t make_desc(int layout, uint32_t smem_addr) {
  const uint32_t lbo = layout == 0 ? 4096 : 16, sbo = layout == 0 ? 128 : 1024;
  const uint64_t type = layout == 0 ? 0 : 2;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (type << 61);
}
__device__ inline uint32_t k_step_bytes(int layout) { return layout == 0 ? 2 * 4096 : 32; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}

// mode: 0 = correctness (4 K-steps, D -> global), >0 = timing with `reps` x 4 K-steps
__global__ void __launch_bounds__(128, 1) k_probe(const float* __restrict__ A, const float* __restrict__ B, int la, int lb, int a_tmem, int n,
                                                  int reps, float* __restrict__ D, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kOpBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kOpBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 2 * kOpBytes + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK; i += 128) *reinterpret_cast<float*>(sA + elem_offset(la, i / kK, i % kK)) = A[i];
  for (int i = tid; i < 256 * kK; i += 128) *reinterpret_cast<float*>(sB + elem_offset(lb, i / kK, i % kK)) = B[(i / kK % 128) * kK + i % kK];
  const uint32_t bar_addr = smem_u32(bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {   // A row `tid` -> TMEM columns [256, 256 + K)
    uint32_t r[32];
    for (int k = 0; k < kK; ++k) r[k] = __float_as_uint(A[tid * kK + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tmem + lane_base + 256),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0 && elect_one()) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint64_t db = make_desc(lb, smem_u32(sB) + ks * k_step_bytes(lb));
        const uint32_t acc = (rep | ks) != 0;
        if (a_tmem) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                       "r"(tmem + 256 + ks * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
        } else {
          const uint64_t da = make_desc(la, smem_u32(sA) + ks * k_step_bytes(la));
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                       "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  }
  __syncwarp();
  mbar_wait(bar_addr, 0);
  t2 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (t0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  if (reps == 1) {
    for (int c = 0; c < 128 / 32; ++c) {
      uint32_t r[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
          "tcgen05.wait::ld.sync.aligned;"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tmem + lane_base + c * 32)
          : "memory");
      for (int e = 0; e < 32; ++e) D[tid * 128 + c * 32 + e] = __uint_as_float(r[e]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  static float hA[kM * kK], hB[128 * kK], hD[kM * 128], ref[kM * 128];
  for (int m = 0; m < kM; ++m)
    for (int k = 0; k < kK; ++k) hA[m * kK + k] = (float)(((m * 3 + k * 5) % 7) - 3);
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < kK; ++k) hB[n * kK + k] = (float)(((n * 2 + k * 7) % 5) - 2);
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < kK; ++k) s += hA[m * kK + k] * hB[n * kK + k];
      ref[m * 128 + n] = s;
    }
  float *dA, *dB, *dD;
  long long* dC;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const int smem = 2 * kOpBytes + 64;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[2] = {"K-major/none", "K-major/sw128"};
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int l = 0; l < 2; ++l) {
      cudaMemset(dD, 0xff, sizeof(hD));
      k_probe<<<1, 128, smem>>>(dA, dB, l, l, a_tmem, 128, 1, dD, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("{\"probe\": \"mma_rate\", \"cuda_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int i = 0; i < kM * 128; ++i) { const double d = fabs((double)hD[i] - (double)ref[i]); if (!(d <= maxerr)) maxerr = d; }
      printf("{\"probe\": \"mma_layout_check\", \"a_src\": \"%s\", \"layout\": \"%s\", \"max_abs_err\": %.3f}\n", a_tmem ? "tmem" : "smem", names[l], maxerr);
      for (int n : {128, 256, 16}) {
        const int reps = 64;
        long long hc[2];
        for (int it = 0; it < 2; ++it) {
          k_probe<<<1, 128, smem>>>(dA, dB, l, l, a_tmem, n, reps, dD, dC);
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("{\"probe\": \"mma_rate\", \"cuda_error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(hc, dC, 16, cudaMemcpyDeviceToHost);
        printf("{\"probe\": \"mma_rate\", \"a_src\": \"%s\", \"layout\": \"%s\", \"N\": %d, \"mmas\": %d, \"issue_cycles_per_mma\": %.1f, \"done_cycles_per_mma\": %.1f, \"floor\": %d}\n",
               a_tmem ? "tmem" : "smem", names[l], n, reps * 4, (double)hc[0] / (reps * 4), (double)hc[1] / (reps * 4), 128 * n / 256);
      }
      fflush(stdout);
    }
  return 0;
}
