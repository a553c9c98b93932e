// march_common.cuh -- ray set-up and contracted sampling shared by the fused march kernels.
#pragma once
#include "trilinear.cuh"

namespace ubn {

struct MarchParams {
  float cx, cy, cz, rx, ry, rz;   // scene center / radius
  float B, A;                     // contraction constants
  int l2norm;
  int S;
  float shift, interval, thres;
  int use_cumdist;
  float cumdist_thres;
  int use_mask;
  int msz[3];
  float mscale[3], mshift[3];
};

inline MarchParams make_params(const UbnMarchCfg* c) {
  MarchParams p;
  p.cx = c->scene_center[0]; p.cy = c->scene_center[1]; p.cz = c->scene_center[2];
  p.rx = c->scene_radius[0]; p.ry = c->scene_radius[1]; p.rz = c->scene_radius[2];
  p.B = c->contract_B; p.A = c->contract_A;
  p.l2norm = c->contracted_norm;
  p.S = c->n_samples;
  p.shift = c->act_shift; p.interval = c->interval; p.thres = c->fast_color_thres;
  p.use_cumdist = c->use_cumdist; p.cumdist_thres = c->cumdist_thres;
  p.use_mask = c->use_maskcache;
  for (int a = 0; a < 3; ++a) { p.msz[a] = c->mask_sz[a]; p.mscale[a] = c->mask_scale[a]; p.mshift[a] = c->mask_shift[a]; }
  return p;
}

struct Ray {
  float ox, oy, oz, dx, dy, dz;   // normalised origin, unit direction
};

// ||v|| exactly as torch's CUDA reduction evaluates x.norm(dim=-1) on 3-vectors: the lanes of the reduced
// dimension are combined by a shuffle tree, i.e. sqrt((x*x + z*z) + y*y) with every product and sum rounded
// separately (probed on B200: 0 mismatches in 2^20 random vectors; the "natural" orders mismatch in 12-15 %).
// The reference runs these norms as torch ops (dcvgo.py:240,253,288; FourierGrid_model.py:523,537), so matching
// them bit-for-bit keeps the threshold decisions downstream (inner mask, cumdist, mask-cache rounding) identical.
__device__ __forceinline__ float norm3_torch(float x, float y, float z) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)), __fmul_rn(y, y)));
}

// rays_o = (o - center) / radius ; rays_d = d / ||d||     (torch elementwise: no fma contraction)
__device__ __forceinline__ Ray load_ray(const float* __restrict__ o, const float* __restrict__ d, const MarchParams& p) {
  Ray r;
  r.ox = __fdiv_rn(__fsub_rn(o[0], p.cx), p.rx);
  r.oy = __fdiv_rn(__fsub_rn(o[1], p.cy), p.ry);
  r.oz = __fdiv_rn(__fsub_rn(o[2], p.cz), p.rz);
  const float n = norm3_torch(d[0], d[1], d[2]);
  r.dx = __fdiv_rn(d[0], n);
  r.dy = __fdiv_rn(d[1], n);
  r.dz = __fdiv_rn(d[2], n);
  return r;
}

// contracted sample position at parameter t; returns inner flag (norm <= 1)
__device__ __forceinline__ bool sample_point(const Ray& r, float t, const MarchParams& p, float& x, float& y, float& z) {
  x = __fadd_rn(r.ox, __fmul_rn(r.dx, t));
  y = __fadd_rn(r.oy, __fmul_rn(r.dy, t));
  z = __fadd_rn(r.oz, __fmul_rn(r.dz, t));
  float n;
  if (p.l2norm) n = norm3_torch(x, y, z);
  else          n = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
  const bool inner = (n <= 1.f);
  if (!inner) {
    // torch evaluates `bg_len / norm` (Python scalar / tensor) as norm.reciprocal() * bg_len -- Tensor.__rtruediv__ --
    // i.e. two roundings; reproduced here so contracted points match the reference's torch ops bit-for-bit
    const float f = __fsub_rn(p.B, __fmul_rn(__frcp_rn(n), p.A));
    x = __fmul_rn(__fdiv_rn(x, n), f);
    y = __fmul_rn(__fdiv_rn(y, n), f);
    z = __fmul_rn(__fdiv_rn(z, n), f);
  }
  return inner;
}

struct CellR {
  int v;            // base voxel index  (x0*Y + y0)*Z + z0, pre-clamped
  float fx, fy, fz; // fractions in [0,1]
};

__device__ __forceinline__ CellR make_cell(float cx, float cy, float cz, int X, int Y, int Z) {
  CellR c;
  const float x0 = fminf(fmaxf(floorf(cx), 0.f), (float)(X - 2));
  const float y0 = fminf(fmaxf(floorf(cy), 0.f), (float)(Y - 2));
  const float z0 = fminf(fmaxf(floorf(cz), 0.f), (float)(Z - 2));
  c.fx = cx - x0; c.fy = cy - y0; c.fz = cz - z0;
  c.v = ((int)x0 * Y + (int)y0) * Z + (int)z0;
  return c;
}

// Visit the kP slabs of a FourierGrid in natural order with their continuous source indices.  Slabs 2k+1 / 2k+2 are sin / cos
// of the SAME argument 2^k x, so one sincosf per (axis, frequency) serves both: half the range reductions of separate sinf / cosf
// calls.  sincosf is bit-identical to the pair on this toolchain for every float |a| <= 8 (scripts/probe_sincos.cu, run on the
// B200: 0 mismatches in 2.18e9 arguments), so every coordinate -- and with it raw_density -- keeps its bits.
template <int kP, typename F>
__device__ __forceinline__ void for_each_slab(const GridView& g, float nx, float ny, float nz, F&& f) {
  f(0, src_index(nx, g.X), src_index(ny, g.Y), src_index(nz, g.Z));
#pragma unroll
  for (int k = 0; k < (kP - 1) / 2; ++k) {
    const float m = (float)(1 << k);
    float sx, cx, sy, cy, sz, cz;
    sincosf(__fmul_rn(m, nx), &sx, &cx);
    sincosf(__fmul_rn(m, ny), &sy, &cy);
    sincosf(__fmul_rn(m, nz), &sz, &cz);
    f(2 * k + 1, src_index(sx, g.X), src_index(sy, g.Y), src_index(sz, g.Z));
    f(2 * k + 2, src_index(cx, g.X), src_index(cy, g.Y), src_index(cz, g.Z));
  }
}

// grid_density_at for a contiguous single-channel grid (sv == 1) with kP slabs known at compile time, 32-bit voxel offsets and
// pre-clamped cells: no per-corner bounds predicates, no 64-bit index arithmetic, no runtime slab loop.  Contracted / Fourier-warped
// coordinates never leave [-1, 1], where the clamped cell gives the same eight (value, weight) pairs in the same order as
// trilerp1 (a corner that trilerp1 skips as out of range has weight exactly 0 here), so the result is bit-identical.
template <int kP>
__device__ __forceinline__ float grid_density_fast(const GridView& g, float x, float y, float z) {
  const float nx = norm_coord(x, g.mn[0], g.len[0]);
  const float ny = norm_coord(y, g.mn[1], g.len[1]);
  const float nz = norm_coord(z, g.mn[2], g.len[2]);
  const int dY = g.Z, dX = g.Y * g.Z;
  SlabMean acc;
  for_each_slab<kP>(g, nx, ny, nz, [&](int s, float cx, float cy, float cz) {
    const CellR c = make_cell(cx, cy, cz, g.X, g.Y, g.Z);
    const float* rec = g.data + (int64_t)s * g.sp + c.v;
    float a = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {          // tnw .. bse, z fastest (ATen's order)
      const int bx = corner >> 2, by = (corner >> 1) & 1, bz = corner & 1;
      const float w = ((bz ? c.fz : 1.f - c.fz) * (by ? c.fy : 1.f - c.fy)) * (bx ? c.fx : 1.f - c.fx);
      a = fmaf(__ldg(rec + bx * dX + by * dY + bz), w, a);
    }
    acc.add(s, a);
  });
  return acc.mean(kP);
}

// adjoint of grid_density_fast: the four (x, y) edges of the cell, two z-adjacent floats each.  An edge whose lower corner is
// 8-byte aligned goes out as one red.v2 {w0, w1}; otherwise as red.v2 {0, w0} at the aligned pair below it (adding +0 to the
// neighbouring voxel: a no-op on the value) plus one scalar red for the upper corner -- the same instruction stream for every
// lane, no divergent alignment branch, 4 + (0..4) reduction instructions per slab instead of 4..12.
template <int kP>
__device__ __forceinline__ void grid_density_scatter_fast(float* __restrict__ grad, const GridView& g, float nx, float ny, float nz,
                                                          float gd) {
  const int dY = g.Z, dX = g.Y * g.Z;
  for_each_slab<kP>(g, nx, ny, nz, [&](int s, float cx, float cy, float cz) {
    const CellR c = make_cell(cx, cy, cz, g.X, g.Y, g.Z);
    float* rec = grad + (int64_t)s * g.sp + c.v;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int bx = e >> 1, by = e & 1;
      const float wxy_lo = (1.f - c.fz) * (by ? c.fy : 1.f - c.fy), wxy_hi = c.fz * (by ? c.fy : 1.f - c.fy);
      const float wx = bx ? c.fx : 1.f - c.fx;
      const float w0 = (wxy_lo * wx) * gd, w1 = (wxy_hi * wx) * gd;
      float* a = rec + bx * dX + by * dY;
      const bool odd = (reinterpret_cast<uintptr_t>(a) & 4) != 0;
      red_add_v2(a - (odd ? 1 : 0), odd ? 0.f : w0, odd ? w0 : w1);
      if (odd) atomicAdd(a + 1, w1);
    }
  });
}

constexpr int kMarchWarps = 4;

}  // namespace ubn
