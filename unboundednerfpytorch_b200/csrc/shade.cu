// shade.cu -- the rgbnet (feature -> RGB MLP) of the hot path as fused, activation-resident kernels.
//
// Reference: rgb = sigmoid(rgbnet(cat[k0, viewdirs_emb[ray_id]])) with rgbnet = Linear(12+27,128)-ReLU-
// Linear(128,128)-ReLU-Linear(128,3) (FourierGrid_model.py:231-242,631-637; dcvgo.py:103-114,337-342), executed as
// ~10 torch kernels that round-trip [M,39], [M,128], [M,128] activations through HBM (about 1.2 KB per sample,
// more than the grid reads themselves -- SURVEY.md 8a a9).  Here:
//   * the view-direction part of layer 1 is constant per ray, so it is hoisted: vb[r] = W1[:,12:] . emb(viewdir_r) + b1
//     (an [N,128] table, computed by the host with one tiny GEMM) and layer 1 becomes K = 12;
//   * one persistent CTA per SM keeps all weights in shared memory and pushes 128-sample tiles through the three
//     layers with the hidden activations living in shared memory / registers only;
//   * fp32 FFMA arithmetic ("exact" mode): results differ from the reference only by fp32 re-association, which is
//     what the 1e-5 parity gate of the north star needs (a single bf16 / tf32 pass cannot meet it);
//   * the training variant also writes the two post-ReLU hidden activations, which the backward kernel consumes
//     instead of recomputing two 128^3-per-tile GEMMs on CUDA cores.
// Backward: dZ2 = (dz3 . W3) * [H2>0]; dW2 += dZ2^T H1; dH1 = dZ2 W2; dZ1 = dH1 * [H1>0]; dW1k += dZ1^T X;
// dX = dZ1 W1k; dvb[ray] += dZ1 (the host turns dvb into dW1[:,12:], db1 with the same tiny GEMM).
#include "common.cuh"

namespace ubn {

constexpr int kW = 128;        // hidden width
constexpr int kF = 12;         // k0 feature channels
constexpr int kFwdTile = 128;  // samples per forward tile
constexpr int kBwdTile = 64;   // samples per backward tile
constexpr int kThreads = 256;

// ---- shared-memory plans (floats) ----------------------------------------------------------------------
struct FwdSmem {
  static constexpr int kH1Stride = kFwdTile + 4;                 // padded: transposed stores hit distinct banks
  static constexpr int oW2 = 0;                                  // [k][perm j]  128 x 128
  static constexpr int oH1 = oW2 + kW * kW;                      // [k][s]       128 x 132
  static constexpr int oW1 = oH1 + kW * kH1Stride;               // [k'][perm j] 12 x 128
  static constexpr int oX = oW1 + kF * kW;                       // [k'][s]      12 x 128
  static constexpr int oW3 = oX + kF * kFwdTile;                 // [c][perm j]  3 x 128
  static constexpr int oB2 = oW3 + 3 * kW;                       // [perm j]
  static constexpr int oRay = oB2 + kW;                          // int[128]
  static constexpr int kFloats = oRay + kFwdTile;
};

// permuted hidden index: thread tx owns hidden units j = jj*16 + tx (jj < 8), stored contiguously at tx*8 + jj
__device__ __forceinline__ int perm_pos(int j) { return (j & 15) * 8 + (j >> 4); }

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

template <bool kSave>
__global__ void __launch_bounds__(kThreads, 1) k_shade_fwd(
    const float* __restrict__ feat, const float* __restrict__ vb, const int64_t* __restrict__ ray_id,
    const float* __restrict__ W1k, const float* __restrict__ W2, const float* __restrict__ b2,
    const float* __restrict__ W3, const float* __restrict__ b3, int64_t n_pts, float* __restrict__ rgb,
    float* __restrict__ h1_out, float* __restrict__ h2_out) {
  extern __shared__ __align__(16) float sm[];
  float* sW2 = sm + FwdSmem::oW2;
  float* sH1 = sm + FwdSmem::oH1;
  float* sW1 = sm + FwdSmem::oW1;
  float* sX = sm + FwdSmem::oX;
  float* sW3 = sm + FwdSmem::oW3;
  float* sB2 = sm + FwdSmem::oB2;
  int* sRay = reinterpret_cast<int*>(sm + FwdSmem::oRay);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // weights -> shared (once per CTA): W2[j][k] -> sW2[k][perm j], W1k[j][k'] -> sW1[k'][perm j], W3[c][j] -> sW3[c][perm j]
  for (int i = tid; i < kW * kW; i += kThreads) {
    const int j = i / kW, k = i % kW;
    sW2[k * kW + perm_pos(j)] = W2[i];
  }
  for (int i = tid; i < kW * kF; i += kThreads) {
    const int j = i / kF, k = i % kF;
    sW1[k * kW + perm_pos(j)] = W1k[i];
  }
  for (int i = tid; i < 3 * kW; i += kThreads) {
    const int c = i / kW, j = i % kW;
    sW3[c * kW + perm_pos(j)] = W3[i];
  }
  if (tid < kW) sB2[perm_pos(tid)] = b2[tid];
  const float b3x = b3[0], b3y = b3[1], b3z = b3[2];
  __syncthreads();

  const int64_t n_tiles = (n_pts + kFwdTile - 1) / kFwdTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kFwdTile;
    const int n_here = (int)min((int64_t)kFwdTile, n_pts - base);
    // X tile transposed: sX[k'][s]; rays
    for (int i = tid; i < kFwdTile * kF; i += kThreads) {
      const int s = i / kF, k = i % kF;
      sX[k * kFwdTile + s] = (s < n_here) ? feat[(base + s) * kF + k] : 0.f;
    }
    if (tid < kFwdTile) sRay[tid] = (tid < n_here) ? (int)ray_id[base + tid] : 0;
    __syncthreads();

    // ---- layer 1: acc[i][jj] = sum_k' X[s0+i][k'] * W1k[j][k'] + vb[ray][j]; ReLU ----
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* vrow = vb + (int64_t)sRay[ty * 8 + i] * kW;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) acc[i][jj] = __ldg(vrow + jj * 16 + tx);
    }
#pragma unroll
    for (int k = 0; k < kF; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(sX + k * kFwdTile + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(sX + k * kFwdTile + ty * 8 + 4);
      const float4 w0 = *reinterpret_cast<const float4*>(sW1 + k * kW + tx * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(sW1 + k * kW + tx * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(a[i], w[jj], acc[i][jj]);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = jj * 16 + tx;
      float h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = fmaxf(acc[i][jj], 0.f);
      *reinterpret_cast<float4*>(sH1 + j * FwdSmem::kH1Stride + ty * 8) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(sH1 + j * FwdSmem::kH1Stride + ty * 8 + 4) = make_float4(h[4], h[5], h[6], h[7]);
      if (kSave) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (ty * 8 + i < n_here) h1_out[(base + ty * 8 + i) * kW + j] = h[i];
      }
    }
    __syncthreads();

    // ---- layer 2: acc[i][jj] = b2[j] + sum_k H1[s][k] * W2[j][k]; ReLU ----
    {
      const float4 c0 = *reinterpret_cast<const float4*>(sB2 + tx * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(sB2 + tx * 8 + 4);
      const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i][jj] = c[jj];
    }
#pragma unroll 4
    for (int k = 0; k < kW; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(sH1 + k * FwdSmem::kH1Stride + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(sH1 + k * FwdSmem::kH1Stride + ty * 8 + 4);
      const float4 w0 = *reinterpret_cast<const float4*>(sW2 + k * kW + tx * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(sW2 + k * kW + tx * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) acc[i][jj] = fmaf(a[i], w[jj], acc[i][jj]);
    }
    // ---- layer 3 + sigmoid: partial over this thread's 8 hidden units, reduced over the 16 tx lanes ----
    float w3[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v0 = *reinterpret_cast<const float4*>(sW3 + c * kW + tx * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(sW3 + c * kW + tx * 8 + 4);
      w3[c][0] = v0.x; w3[c][1] = v0.y; w3[c][2] = v0.z; w3[c][3] = v0.w;
      w3[c][4] = v1.x; w3[c][5] = v1.y; w3[c][6] = v1.z; w3[c][7] = v1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float h = fmaxf(acc[i][jj], 0.f);
        if (kSave && ty * 8 + i < n_here) h2_out[(base + ty * 8 + i) * kW + jj * 16 + tx] = h;
        p0 = fmaf(h, w3[0][jj], p0);
        p1 = fmaf(h, w3[1][jj], p1);
        p2 = fmaf(h, w3[2][jj], p2);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        p0 += __shfl_xor_sync(0xffffffffu, p0, o);
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        p2 += __shfl_xor_sync(0xffffffffu, p2, o);
      }
      if (tx == 0 && ty * 8 + i < n_here) {
        float* o = rgb + (base + ty * 8 + i) * 3;
        o[0] = sigmoidf_exact(p0 + b3x);
        o[1] = sigmoidf_exact(p1 + b3y);
        o[2] = sigmoidf_exact(p2 + b3z);
      }
    }
    __syncthreads();   // sH1 / sX / sRay are rewritten by the next tile
  }
}

// ---- backward ------------------------------------------------------------------------------------------
struct BwdSmem {
  static constexpr int kRow = kW + 4;          // 132: padded row of a [sample][hidden] tile
  static constexpr int kTRow = kBwdTile + 4;   // 68 : padded row of a [hidden][sample] tile
  static constexpr int oW2 = 0;                              // W2[j][k] natural          128 x 128
  static constexpr int oH1 = oW2 + kW * kW;                  // H1s[s][k]                 64 x 132
  static constexpr int oDZ = oH1 + kBwdTile * kRow;          // dZ2s[s][j] then dZ1s[s][k] 64 x 132
  static constexpr int oDZT = oDZ + kBwdTile * kRow;         // dZ2t[j][s]                128 x 68
  static constexpr int oX = oDZT + kW * kTRow;               // Xs[s][k']                 64 x 12
  static constexpr int oW1 = oX + kBwdTile * kF;             // W1k[j][k']                128 x 12
  static constexpr int oW3 = oW1 + kW * kF;                  // W3[c][j]                  3 x 128
  static constexpr int oDz3 = oW3 + 3 * kW;                  // dz3[s][4]                 64 x 4
  static constexpr int oRay = oDz3 + kBwdTile * 4;           // int[64]
  static constexpr int kFloats = oRay + kBwdTile;
};

__global__ void __launch_bounds__(kThreads, 1) k_shade_bwd(
    const float* __restrict__ feat, const int64_t* __restrict__ ray_id, const float* __restrict__ W1k,
    const float* __restrict__ W2, const float* __restrict__ W3, const float* __restrict__ rgb,
    const float* __restrict__ h1, const float* __restrict__ h2, const float* __restrict__ g_rgb, int64_t n_pts,
    float* __restrict__ g_feat, float* __restrict__ g_vb, float* __restrict__ gW1k, float* __restrict__ gW2,
    float* __restrict__ gb2, float* __restrict__ gW3, float* __restrict__ gb3) {
  extern __shared__ __align__(16) float sm[];
  float* sW2 = sm + BwdSmem::oW2;
  float* sH1 = sm + BwdSmem::oH1;
  float* sDZ = sm + BwdSmem::oDZ;
  float* sDZT = sm + BwdSmem::oDZT;
  float* sX = sm + BwdSmem::oX;
  float* sW1 = sm + BwdSmem::oW1;
  float* sW3 = sm + BwdSmem::oW3;
  float* sDz3 = sm + BwdSmem::oDz3;
  int* sRay = reinterpret_cast<int*>(sm + BwdSmem::oRay);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int R = BwdSmem::kRow, TR = BwdSmem::kTRow;

  for (int i = tid; i < kW * kW; i += kThreads) sW2[i] = W2[i];
  for (int i = tid; i < kW * kF; i += kThreads) sW1[i] = W1k[i];
  for (int i = tid; i < 3 * kW; i += kThreads) sW3[i] = W3[i];
  __syncthreads();

  // persistent per-thread accumulators
  float aW2[8][8];      // dW2[j = ty*8+a][k = tx*8+b]
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) aW2[a][b] = 0.f;
  float aB2[8], aW3[3][8];   // hidden units j = jj*16 + tx (mapping A), partial over this thread's samples
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) { aB2[jj] = 0.f; aW3[0][jj] = aW3[1][jj] = aW3[2][jj] = 0.f; }
  float aW1[6];         // dW1k[j = tid%128][k' = 6*(tid/128) + q]
#pragma unroll
  for (int q = 0; q < 6; ++q) aW1[q] = 0.f;
  float aB3[3] = {0.f, 0.f, 0.f};   // thread tid < 64 : sample-partial of db3
  // dvb[ray][k = tx*8+q]: running per-ray sum of this thread's dZ1 rows; rays are sorted and every CTA walks a
  // CONTIGUOUS range of tiles, so a run is flushed (8 atomics) only when the ray changes
  float run[8];
  int run_ray = -1;
#pragma unroll
  for (int q = 0; q < 8; ++q) run[q] = 0.f;

  const int64_t n_tiles = (n_pts + kBwdTile - 1) / kBwdTile;
  const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_end = min(n_tiles, (int64_t)(blockIdx.x + 1) * per_cta);
  for (int64_t tile = (int64_t)blockIdx.x * per_cta; tile < tile_end; ++tile) {
    const int64_t base = tile * kBwdTile;
    const int n_here = (int)min((int64_t)kBwdTile, n_pts - base);
    // ---- stage 0: tiles -> shared ----
    for (int i = tid; i < kBwdTile * (kW / 4); i += kThreads) {
      const int s = i / (kW / 4), q = i % (kW / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (s < n_here) v = *reinterpret_cast<const float4*>(h1 + (base + s) * kW + q * 4);
      *reinterpret_cast<float4*>(sH1 + s * R + q * 4) = v;
    }
    for (int i = tid; i < kBwdTile * kF; i += kThreads) {
      const int s = i / kF;
      sX[i] = (s < n_here) ? feat[base * kF + i] : 0.f;
    }
    if (tid < kBwdTile) {
      const int s = tid;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      int r = 0;
      if (s < n_here) {
        const float* o = rgb + (base + s) * 3;
        const float* g = g_rgb + (base + s) * 3;
        d0 = g[0] * (o[0] * (1.f - o[0]));
        d1 = g[1] * (o[1] * (1.f - o[1]));
        d2 = g[2] * (o[2] * (1.f - o[2]));
        r = (int)ray_id[base + s];
      }
      sDz3[s * 4] = d0; sDz3[s * 4 + 1] = d1; sDz3[s * 4 + 2] = d2; sDz3[s * 4 + 3] = 0.f;
      sRay[s] = r;
      aB3[0] += d0; aB3[1] += d1; aB3[2] += d2;
    }
    __syncthreads();

    // ---- stage 1 (mapping A: j = jj*16+tx, s = ty*4+i): dZ2 = (dz3 . W3) * [H2 > 0]; db2, dW3 partials ----
    {
      float w3[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) w3[c][jj] = sW3[c * kW + jj * 16 + tx];
      float dz[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int s = ty * 4 + i;
        const float4 d = *reinterpret_cast<const float4*>(sDz3 + s * 4);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float hv = (s < n_here) ? __ldg(h2 + (base + s) * kW + jj * 16 + tx) : 0.f;
          const float dh = fmaf(d.z, w3[2][jj], fmaf(d.y, w3[1][jj], d.x * w3[0][jj]));
          const float v = hv > 0.f ? dh : 0.f;
          dz[i][jj] = v;
          aB2[jj] += v;
          aW3[0][jj] = fmaf(d.x, hv, aW3[0][jj]);
          aW3[1][jj] = fmaf(d.y, hv, aW3[1][jj]);
          aW3[2][jj] = fmaf(d.z, hv, aW3[2][jj]);
          sDZ[s * R + jj * 16 + tx] = v;
        }
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        *reinterpret_cast<float4*>(sDZT + (jj * 16 + tx) * TR + ty * 4) = make_float4(dz[0][jj], dz[1][jj], dz[2][jj], dz[3][jj]);
    }
    __syncthreads();

    // ---- stage 2 (mapping B: j = ty*8+a, k = tx*8+b): dW2[j][k] += sum_s dZ2[s][j] * H1[s][k] ----
#pragma unroll 2
    for (int s = 0; s < kBwdTile; ++s) {
      const float4 a0 = *reinterpret_cast<const float4*>(sDZ + s * R + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(sDZ + s * R + ty * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sH1 + s * R + tx * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(sH1 + s * R + tx * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) aW2[p][q] = fmaf(a[p], b[q], aW2[p][q]);
    }
    // ---- stage 3 (mapping C: s = ty*4+i, k = tx*8+b): dH1[s][k] = sum_j dZ2t[j][s] * W2[j][k] ----
    float dh[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int b = 0; b < 8; ++b) dh[i][b] = 0.f;
#pragma unroll 4
    for (int j = 0; j < kW; ++j) {
      const float4 a0 = *reinterpret_cast<const float4*>(sDZT + j * TR + ty * 4);
      const float4 b0 = *reinterpret_cast<const float4*>(sW2 + j * kW + tx * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(sW2 + j * kW + tx * 8 + 4);
      const float a[4] = {a0.x, a0.y, a0.z, a0.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) dh[i][q] = fmaf(a[i], b[q], dh[i][q]);
    }
    __syncthreads();   // everybody finished reading sDZ (dZ2s) in stage 2 before it becomes dZ1s
    // dZ1 = dH1 * [H1 > 0] -> sDZ (as dZ1s[s][k]); per-ray accumulation into g_vb
    {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int s = ty * 4 + i;
        const float4 h0 = *reinterpret_cast<const float4*>(sH1 + s * R + tx * 8);
        const float4 h1v = *reinterpret_cast<const float4*>(sH1 + s * R + tx * 8 + 4);
        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1v.x, h1v.y, h1v.z, h1v.w};
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = hv[q] > 0.f ? dh[i][q] : 0.f;
        *reinterpret_cast<float4*>(sDZ + s * R + tx * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(sDZ + s * R + tx * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        if (s < n_here) {
          const int r = sRay[s];
          if (r != run_ray) {
            if (run_ray >= 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q) atomicAdd(g_vb + (int64_t)run_ray * kW + tx * 8 + q, run[q]);
            }
            run_ray = r;
#pragma unroll
            for (int q = 0; q < 8; ++q) run[q] = 0.f;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) run[q] += v[q];
        }
      }
    }
    __syncthreads();
    // ---- stage 4: dW1k[j][k'] += sum_s dZ1[s][j] X[s][k'];  dX[s][k'] = sum_j dZ1[s][j] W1k[j][k'] ----
    {
      const int j = tid & 127, k0 = 6 * (tid >> 7);
#pragma unroll 4
      for (int s = 0; s < kBwdTile; ++s) {
        const float d = sDZ[s * R + j];
#pragma unroll
        for (int q = 0; q < 6; ++q) aW1[q] = fmaf(d, sX[s * kF + k0 + q], aW1[q]);
      }
      const int s = tid >> 2, kk = 3 * (tid & 3);
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
#pragma unroll 4
      for (int jx = 0; jx < kW; ++jx) {
        const float d = sDZ[s * R + jx];
        x0 = fmaf(d, sW1[jx * kF + kk], x0);
        x1 = fmaf(d, sW1[jx * kF + kk + 1], x1);
        x2 = fmaf(d, sW1[jx * kF + kk + 2], x2);
      }
      if (s < n_here) {
        float* o = g_feat + (base + s) * kF + kk;
        o[0] = x0; o[1] = x1; o[2] = x2;
      }
    }
    __syncthreads();
  }

  // ---- flush the per-CTA partial sums (buffers are zero-initialised by the host) ----
  if (run_ray >= 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(g_vb + (int64_t)run_ray * kW + tx * 8 + q, run[q]);
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) atomicAdd(gW2 + (ty * 8 + a) * kW + tx * 8 + b, aW2[a][b]);
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = jj * 16 + tx;
    atomicAdd(gb2 + j, aB2[jj]);
    atomicAdd(gW3 + j, aW3[0][jj]);
    atomicAdd(gW3 + kW + j, aW3[1][jj]);
    atomicAdd(gW3 + 2 * kW + j, aW3[2][jj]);
  }
  {
    const int j = tid & 127, k0 = 6 * (tid >> 7);
#pragma unroll
    for (int q = 0; q < 6; ++q) atomicAdd(gW1k + j * kF + k0 + q, aW1[q]);
  }
  if (tid < kBwdTile) {
    atomicAdd(gb3, aB3[0]); atomicAdd(gb3 + 1, aB3[1]); atomicAdd(gb3 + 2, aB3[2]);
  }
}


// ---- small gradients for the tensor-core backward (shade_tc.cu) ------------------------------------------------------
// k_shade_bwd_tc produces dZ1 [M,128] (HBM) and dW2; everything else is thin and memory-bound and is finished here on
// the CUDA cores from dZ1, H2, X and dz3:  db2, dW3, db3 (from dZ2 recomputed on the fly), dvb[ray] (per-ray sums of
// dZ1), dW1k = dZ1^T X, dX = dZ1 W1k.
struct SmallSmem {
  static constexpr int kRow = kW + 4;
  static constexpr int oDZ = 0;                               // dZ1s[s][k]  64 x 132
  static constexpr int oX = oDZ + kBwdTile * kRow;            // Xs[s][k']   64 x 12
  static constexpr int oW1 = oX + kBwdTile * kF;              // W1k[j][k']  128 x 12
  static constexpr int oW3 = oW1 + kW * kF;                   // W3[c][j]
  static constexpr int oDz3 = oW3 + 3 * kW;
  static constexpr int oRay = oDz3 + kBwdTile * 4;
  static constexpr int kFloats = oRay + kBwdTile;
};

__global__ void __launch_bounds__(kThreads, 2) k_shade_bwd_small(
    const float* __restrict__ feat, const int64_t* __restrict__ ray_id, const float* __restrict__ W1k,
    const float* __restrict__ W3, const float* __restrict__ rgb, const float* __restrict__ h2,
    const float* __restrict__ g_rgb, const float* __restrict__ dz1, int64_t n_pts, float* __restrict__ g_feat,
    float* __restrict__ g_vb, float* __restrict__ gW1k, float* __restrict__ gb2, float* __restrict__ gW3,
    float* __restrict__ gb3) {
  extern __shared__ __align__(16) float sm[];
  float* sDZ = sm + SmallSmem::oDZ;
  float* sX = sm + SmallSmem::oX;
  float* sW1 = sm + SmallSmem::oW1;
  float* sW3 = sm + SmallSmem::oW3;
  float* sDz3 = sm + SmallSmem::oDz3;
  int* sRay = reinterpret_cast<int*>(sm + SmallSmem::oRay);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int R = SmallSmem::kRow;
  for (int i = tid; i < kW * kF; i += kThreads) sW1[i] = W1k[i];
  for (int i = tid; i < 3 * kW; i += kThreads) sW3[i] = W3[i];
  __syncthreads();
  float w3[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) w3[c][jj] = sW3[c * kW + jj * 16 + tx];
  float aB2[8], aW3[3][8], aW1[6], aB3[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) { aB2[jj] = 0.f; aW3[0][jj] = aW3[1][jj] = aW3[2][jj] = 0.f; }
#pragma unroll
  for (int q = 0; q < 6; ++q) aW1[q] = 0.f;
  float run[8];
  int run_ray = -1;
#pragma unroll
  for (int q = 0; q < 8; ++q) run[q] = 0.f;

  const int64_t n_tiles = (n_pts + kBwdTile - 1) / kBwdTile;
  const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_end = min(n_tiles, (int64_t)(blockIdx.x + 1) * per_cta);
  for (int64_t tile = (int64_t)blockIdx.x * per_cta; tile < tile_end; ++tile) {
    const int64_t base = tile * kBwdTile;
    const int n_here = (int)min((int64_t)kBwdTile, n_pts - base);
    for (int i = tid; i < kBwdTile * (kW / 4); i += kThreads) {
      const int s = i / (kW / 4), q = i % (kW / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (s < n_here) v = __ldg(reinterpret_cast<const float4*>(dz1 + (base + s) * kW + q * 4));
      *reinterpret_cast<float4*>(sDZ + s * R + q * 4) = v;
    }
    for (int i = tid; i < kBwdTile * kF; i += kThreads) {
      const int s = i / kF;
      sX[i] = (s < n_here) ? feat[base * kF + i] : 0.f;
    }
    if (tid < kBwdTile) {
      const int s = tid;
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      int r = 0;
      if (s < n_here) {
        const float* o = rgb + (base + s) * 3;
        const float* g = g_rgb + (base + s) * 3;
        d0 = g[0] * (o[0] * (1.f - o[0]));
        d1 = g[1] * (o[1] * (1.f - o[1]));
        d2 = g[2] * (o[2] * (1.f - o[2]));
        r = (int)ray_id[base + s];
      }
      sDz3[s * 4] = d0; sDz3[s * 4 + 1] = d1; sDz3[s * 4 + 2] = d2; sDz3[s * 4 + 3] = 0.f;
      sRay[s] = r;
      aB3[0] += d0; aB3[1] += d1; aB3[2] += d2;
    }
    __syncthreads();
    // db2 / dW3 (hidden units j = jj*16 + tx, samples s = ty*4 + i)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = ty * 4 + i;
      const float4 d = *reinterpret_cast<const float4*>(sDz3 + s * 4);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float hv = (s < n_here) ? __ldg(h2 + (base + s) * kW + jj * 16 + tx) : 0.f;
        const float dh = fmaf(d.z, w3[2][jj], fmaf(d.y, w3[1][jj], d.x * w3[0][jj]));
        aB2[jj] += hv > 0.f ? dh : 0.f;
        aW3[0][jj] = fmaf(d.x, hv, aW3[0][jj]);
        aW3[1][jj] = fmaf(d.y, hv, aW3[1][jj]);
        aW3[2][jj] = fmaf(d.z, hv, aW3[2][jj]);
      }
    }
    // dvb: per-ray running sums of this thread's dZ1 rows (k = tx*8 + q)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int s = ty * 4 + i;
      if (s < n_here) {
        const float4 v0 = *reinterpret_cast<const float4*>(sDZ + s * R + tx * 8);
        const float4 v1 = *reinterpret_cast<const float4*>(sDZ + s * R + tx * 8 + 4);
        const int r = sRay[s];
        if (r != run_ray) {
          if (run_ray >= 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(g_vb + (int64_t)run_ray * kW + tx * 8 + q, run[q]);
          }
          run_ray = r;
#pragma unroll
          for (int q = 0; q < 8; ++q) run[q] = 0.f;
        }
        run[0] += v0.x; run[1] += v0.y; run[2] += v0.z; run[3] += v0.w;
        run[4] += v1.x; run[5] += v1.y; run[6] += v1.z; run[7] += v1.w;
      }
    }
    // dW1k, dX
    {
      const int j = tid & 127, k0 = 6 * (tid >> 7);
#pragma unroll 4
      for (int s = 0; s < kBwdTile; ++s) {
        const float d = sDZ[s * R + j];
#pragma unroll
        for (int q = 0; q < 6; ++q) aW1[q] = fmaf(d, sX[s * kF + k0 + q], aW1[q]);
      }
      const int s = tid >> 2, kk = 3 * (tid & 3);
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
#pragma unroll 4
      for (int jx = 0; jx < kW; ++jx) {
        const float d = sDZ[s * R + jx];
        x0 = fmaf(d, sW1[jx * kF + kk], x0);
        x1 = fmaf(d, sW1[jx * kF + kk + 1], x1);
        x2 = fmaf(d, sW1[jx * kF + kk + 2], x2);
      }
      if (s < n_here) {
        float* o = g_feat + (base + s) * kF + kk;
        o[0] = x0; o[1] = x1; o[2] = x2;
      }
    }
    __syncthreads();
  }
  if (run_ray >= 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(g_vb + (int64_t)run_ray * kW + tx * 8 + q, run[q]);
  }
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int j = jj * 16 + tx;
    atomicAdd(gb2 + j, aB2[jj]);
    atomicAdd(gW3 + j, aW3[0][jj]);
    atomicAdd(gW3 + kW + j, aW3[1][jj]);
    atomicAdd(gW3 + 2 * kW + j, aW3[2][jj]);
  }
  {
    const int j = tid & 127, k0 = 6 * (tid >> 7);
#pragma unroll
    for (int q = 0; q < 6; ++q) atomicAdd(gW1k + j * kF + k0 + q, aW1[q]);
  }
  if (tid < kBwdTile) {
    atomicAdd(gb3, aB3[0]); atomicAdd(gb3 + 1, aB3[1]); atomicAdd(gb3 + 2, aB3[2]);
  }
}

static int shade_grid() { return kNumSMs; }

}  // namespace ubn

using namespace ubn;

extern "C" {

int ubn_rgbnet_fwd(const float* feat, const float* view_bias, const int64_t* ray_id, const float* W1k, const float* W2,
                   const float* b2, const float* W3, const float* b3, int64_t n_pts, float* rgb, float* h1_save,
                   float* h2_save, void* stream) {
  if (n_pts <= 0) return 0;
  const size_t smem = sizeof(float) * FwdSmem::kFloats;
  const bool save = h1_save != nullptr && h2_save != nullptr;
  cudaError_t e;
  if (save) {
    e = cudaFuncSetAttribute(k_shade_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return finish(e);
    k_shade_fwd<true><<<shade_grid(), kThreads, smem, as_stream(stream)>>>(feat, view_bias, ray_id, W1k, W2, b2, W3, b3,
                                                                          n_pts, rgb, h1_save, h2_save);
  } else {
    e = cudaFuncSetAttribute(k_shade_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return finish(e);
    k_shade_fwd<false><<<shade_grid(), kThreads, smem, as_stream(stream)>>>(feat, view_bias, ray_id, W1k, W2, b2, W3, b3,
                                                                           n_pts, rgb, nullptr, nullptr);
  }
  UBN_LAUNCH_CHECK();
  return 0;
}

int ubn_rgbnet_bwd(const float* feat, const int64_t* ray_id, const float* W1k, const float* W2, const float* W3,
                   const float* rgb, const float* h1_save, const float* h2_save, const float* grad_rgb, int64_t n_pts,
                   float* grad_feat, float* grad_view_bias, float* grad_W1k, float* grad_W2, float* grad_b2,
                   float* grad_W3, float* grad_b3, void* stream) {
  if (n_pts <= 0) return 0;
  const size_t smem = sizeof(float) * BwdSmem::kFloats;
  cudaError_t e = cudaFuncSetAttribute(k_shade_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return finish(e);
  k_shade_bwd<<<shade_grid(), kThreads, smem, as_stream(stream)>>>(feat, ray_id, W1k, W2, W3, rgb, h1_save, h2_save, grad_rgb,
                                                                  n_pts, grad_feat, grad_view_bias, grad_W1k, grad_W2,
                                                                  grad_b2, grad_W3, grad_b3);
  UBN_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"


extern "C" int ubn_rgbnet_bwd_small(const float* feat, const int64_t* ray_id, const float* W1k, const float* W3,
                                    const float* rgb, const float* h2_save, const float* grad_rgb, const float* dz1,
                                    int64_t n_pts, float* grad_feat, float* grad_view_bias, float* grad_W1k, float* grad_b2,
                                    float* grad_W3, float* grad_b3, void* stream) {
  if (n_pts <= 0) return 0;
  const size_t smem = sizeof(float) * SmallSmem::kFloats;
  cudaError_t e = cudaFuncSetAttribute(k_shade_bwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return finish(e);
  k_shade_bwd_small<<<2 * kNumSMs, kThreads, smem, as_stream(stream)>>>(feat, ray_id, W1k, W3, rgb, h2_save, grad_rgb, dz1, n_pts,
                                                                        grad_feat, grad_view_bias, grad_W1k, grad_b2, grad_W3,
                                                                        grad_b3);
  UBN_LAUNCH_CHECK();
  return 0;
}
