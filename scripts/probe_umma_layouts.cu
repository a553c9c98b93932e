// probe_umma_layouts.cu -- hardware probe (B200, sm_100a): which shared-memory descriptor encodings does tcgen05.mma kind::tf32
// accept for MN-major operands?  Prints, per (A layout, B layout, A source) combination, the max abs error of
// D[128x128] = A[128xK] . B[128xK]^T (K = 32, small-integer inputs => exact in tf32 / fp32) against a host reference.
//
// Layout codes:  0 = K-major, no swizzle (the form shade_tc.cu already uses; known good)
//                1 = MN-major, no swizzle, descriptor as CUTLASS builds it   (LBO field = 8-K-group stride, SBO field = 16-byte MN-chunk stride)
//                2 = MN-major, no swizzle, LBO / SBO fields swapped
//                3 = MN-major, 128-byte swizzle, CUTLASS form                (LBO field = 32-element MN-group stride, SBO field = 8-K-group stride)
//                4 = MN-major, 128-byte swizzle, LBO / SBO fields swapped
//                5 = MN-major, SWIZZLE_128B_BASE32B (layout type 1: the only MN-major form CUTLASS builds for 32-bit operands), CUTLASS form
//                6 = same, LBO / SBO fields swapped
//                7 = K-major, no swizzle, K-panel stride (LBO) padded by 16 bytes (bank-conflict-free transposing 4-byte stores)
// A source: 0 = shared memory (SS form), 1 = tensor memory (TS form; A is then always "row = lane", layout code ignored).
//
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a scripts/probe_umma_layouts.cu -o scripts/bin/probe_umma_layouts
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kM = 128, kN = 128, kK = 32;
constexpr uint32_t kOpBytes = 20 * 1024;        // >= 16 KB per operand (+ room for the padded-LBO variant), keeps 1024-byte alignment

__host__ __device__ inline uint32_t elem_offset(int layout, int r, int k) {
  if (layout == 0) return (uint32_t)(k >> 2) * 2048u + (uint32_t)r * 16u + (uint32_t)(k & 3) * 4u;
  if (layout == 1 || layout == 2) return (uint32_t)(r >> 2) * 128u + (uint32_t)(r & 3) * 4u + (uint32_t)(k & 7) * 16u + (uint32_t)(k >> 3) * 4096u;
  if (layout == 3 || layout == 4) {
    uint32_t off = (uint32_t)(r >> 5) * 4096u + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u + (uint32_t)(r & 31) * 4u;
    return off ^ (((off >> 7) & 7u) << 4);
  }
  if (layout == 5 || layout == 6) {   // MN-major SWIZZLE_128B_BASE32B: 32 MN x 4 K atoms (512 B), byte bits [2,4) ^= bits [4,6)
    uint32_t off = (uint32_t)(r >> 5) * 4096u + (uint32_t)(k >> 2) * 512u + (uint32_t)(k & 3) * 128u + (uint32_t)(r & 31) * 4u;
    return off ^ (((off >> 4) & 3u) << 2);
  }
  return (uint32_t)(k >> 2) * 2064u + (uint32_t)r * 16u + (uint32_t)(k & 3) * 4u;    // 7: K-major, K-panel stride padded by 16 B
}

__device__ inline uint64_t make_desc(int layout, uint32_t smem_addr) {
  uint32_t lbo, sbo;
  uint64_t type = 0;
  switch (layout) {
    case 0: lbo = 2048; sbo = 128; break;
    case 1: lbo = 4096; sbo = 128; break;
    case 2: lbo = 128; sbo = 4096; break;
    case 3: lbo = 4096; sbo = 1024; type = 2; break;
    case 4: lbo = 1024; sbo = 4096; type = 2; break;
    case 5: lbo = 4096; sbo = 512; type = 1; break;
    case 6: lbo = 512; sbo = 4096; type = 1; break;
    default: lbo = 2064; sbo = 128; break;
  }
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (type << 61);
}

__device__ inline uint32_t k_step_bytes(int layout) {       // start-address advance per K = 8 step
  if (layout == 0) return 4096;
  if (layout == 1 || layout == 2) return 4096;
  if (layout == 3 || layout == 4) return 1024;
  if (layout == 5 || layout == 6) return 1024;     // two 4-K atoms
  return 2 * 2064;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_probe(const float* __restrict__ A, const float* __restrict__ B, int la, int lb, int a_tmem,
                                                  float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kOpBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kOpBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 2 * kOpBytes + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kM * kK; i += 128) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<float*>(sA + elem_offset(la, r, k)) = A[i];
    *reinterpret_cast<float*>(sB + elem_offset(lb, r, k)) = B[i];
  }
  const uint32_t bar_addr = smem_u32(bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (a_tmem) {   // A row `tid` -> TMEM columns [128, 128 + K)
    uint32_t r[32];
    for (int k = 0; k < kK; ++k) r[k] = __float_as_uint(A[tid * kK + k]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tmem + lane_base + 128),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((la != 0 && la != 7 && !a_tmem) ? (1u << 15) : 0u) | ((lb != 0 && lb != 7) ? (1u << 16) : 0u) |
                           ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
    for (int ks = 0; ks < kK / 8; ++ks) {
      const uint64_t db = make_desc(lb, smem_u32(sB) + ks * k_step_bytes(lb));
      const uint32_t acc = ks > 0;
      if (a_tmem) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
                     "r"(tmem + 128 + ks * 8), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
      } else {
        const uint64_t da = make_desc(la, smem_u32(sA) + ks * k_step_bytes(la));
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar_addr),
      "r"(0u)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < kN / 32; ++c) {
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
    for (int e = 0; e < 32; ++e) D[tid * kN + c * 32 + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int only = argc == 4 ? 1 : 0;
  const int o_t = only ? atoi(argv[1]) : 0, o_a = only ? atoi(argv[2]) : 0, o_b = only ? atoi(argv[3]) : 0;
  static float hA[kM * kK], hB[kN * kK], hD[kM * kN], ref[kM * kN];
  for (int m = 0; m < kM; ++m)
    for (int k = 0; k < kK; ++k) hA[m * kK + k] = (float)(((m * 3 + k * 5) % 7) - 3);
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kK; ++k) hB[n * kK + k] = (float)(((n * 2 + k * 7) % 5) - 2);
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      float s = 0;
      for (int k = 0; k < kK; ++k) s += hA[m * kK + k] * hB[n * kK + k];
      ref[m * kN + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  const int smem = 2 * kOpBytes + 64;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[8] = {"K-major/none", "MN/none/cutlass", "MN/none/swapped", "MN/sw128/cutlass", "MN/sw128/swapped",
                          "MN/sw128_base32b/cutlass", "MN/sw128_base32b/swapped", "K-major/none/padded-LBO"};
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int la = 0; la < (a_tmem ? 1 : 8); ++la)
      for (int lb = 0; lb < 8; ++lb) {
        if (only && (a_tmem != o_t || la != o_a || lb != o_b)) continue;
        cudaMemset(dD, 0xff, sizeof(hD));
        k_probe<<<1, 128, smem>>>(dA, dB, la, lb, a_tmem, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("{\"probe\": \"umma_layout\", \"a_src\": \"%s\", \"A\": \"%s\", \"B\": \"%s\", \"cuda_error\": \"%s\"}\n", a_tmem ? "tmem" : "smem",
                 names[la], names[lb], cudaGetErrorString(e));
          return 1;      // a sticky error poisons the context: stop
        }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double maxerr = 0;
        int zeros = 0, nans = 0;
        for (int i = 0; i < kM * kN; ++i) {
          if (hD[i] != hD[i]) { ++nans; continue; }
          const double d = fabs((double)hD[i] - (double)ref[i]);
          if (d > maxerr) maxerr = d;
          if (hD[i] == 0.f) ++zeros;
        }
        printf("{\"probe\": \"umma_layout\", \"a_src\": \"%s\", \"A\": \"%s\", \"B\": \"%s\", \"max_abs_err\": %.3f, \"zeros\": %d, \"nans\": %d, \"ok\": %s}\n",
               a_tmem ? "tmem" : "smem", names[la], names[lb], maxerr, zeros, nans, (maxerr == 0 && nans == 0) ? "true" : "false");
        fflush(stdout);
      }
  return 0;
}
