"""Fused ray march (autograd.Function) over the C ABI: the hot path of FourierGridModel.forward
(FourierGrid_model.py:554-621) and DirectContractedVoxGO.forward (dcvgo.py:264-331) up to and including
the feature-grid read, in 3 launches forward (pass A, scan, pass B) and 2 backward.
"""
import functools
import os

import numpy as np
import torch

from . import _cabi, ops
from ._cabi import UbnMarchCfg, c_i64, check, ptr, stream_of
from .grid import grid_desc


@functools.lru_cache(maxsize=64)
def _t_schedule_cpu(world_len, stepsize, bg_len, t_boundary):
    """t table exactly as the reference builds it (torch.linspace on fp32, midpoints):
    dcvgo.py:241-248 (t_boundary = 2) and FourierGrid_model.py:524-532 (t_boundary = 1.5)."""
    n_inner = int(2 / (2 + 2 * bg_len) * world_len / stepsize) + 1
    n_outer = n_inner
    b_inner = torch.linspace(0, t_boundary, n_inner + 1, dtype=torch.float32, device='cpu')
    b_outer = t_boundary / torch.linspace(1, 1 / 128, n_outer + 1, dtype=torch.float32, device='cpu')
    return torch.cat([(b_inner[1:] + b_inner[:-1]) * 0.5, (b_outer[1:] + b_outer[:-1]) * 0.5]).contiguous()


_t_dev_cache = {}


def t_schedule(world_len, stepsize, bg_len, t_boundary, device):
    key = (int(world_len), float(stepsize), float(bg_len), float(t_boundary), str(device))
    hit = _t_dev_cache.get(key)
    if hit is None:
        hit = _t_schedule_cpu(*key[:4]).to(device)
        _t_dev_cache[key] = hit
    return hit


TMA_STATS = None          # optional torch.int64[2] CUDA tensor: += {blocks served by TMA, blocks served by the fallback} (tests / bench)


def tma_supported(k0_grid):
    """Single-slab 12-channel channels-last feature grid (DenseGrid k0 of the DCVGO / DVGO family)."""
    return (k0_grid.is_cuda and k0_grid.dim() == 5 and k0_grid.shape[0] == 1 and k0_grid.shape[1] == 12 and k0_grid.stride(1) == 1
            and min(k0_grid.shape[2:]) >= 2)


def make_cfg(scene_center, scene_radius, bg_len, contracted_norm, n_samples, act_shift, interval,
             fast_color_thres, cumdist_thres=None, mask=None, mask_scale=None, mask_shift=None):
    c = UbnMarchCfg()
    for a in range(3):
        c.scene_center[a] = float(scene_center[a])
        c.scene_radius[a] = float(scene_radius[a])
    # torch narrows the Python doubles (1+bg_len) and bg_len to fp32 when they meet an fp32 tensor
    c.contract_B = float(np.float32(1 + bg_len))
    c.contract_A = float(np.float32((1 + bg_len) * 1.0 - 1.0))
    if contracted_norm == 'inf':
        c.contracted_norm = 0
    elif contracted_norm == 'l2':
        c.contracted_norm = 1
    else:
        raise NotImplementedError(contracted_norm)
    c.n_samples = int(n_samples)
    c.act_shift = float(act_shift)
    c.interval = float(interval)
    c.fast_color_thres = float(fast_color_thres)
    c.use_cumdist = 1 if cumdist_thres is not None else 0
    c.cumdist_thres = float(cumdist_thres) if cumdist_thres is not None else 0.0
    c.use_maskcache = 1 if mask is not None else 0
    if mask is not None:
        for a in range(3):
            c.mask_sz[a] = int(mask.shape[a])
            c.mask_scale[a] = float(mask_scale[a])
            c.mask_shift[a] = float(mask_shift[a])
    return c


class March(torch.autograd.Function):
    """(density_grid, k0_grid, rays) -> compacted per-survivor records.

    Returns (weights[M], alphainv_last[N], raw_alpha[M], raw_density[M], k0_feat[M,C], ray_id[M] i64,
    step_id[M] i64, t[M], inner[M] bool).  Differentiable wrt density_grid and k0_grid through weights,
    alphainv_last, raw_alpha, raw_density and k0_feat.
    """

    @staticmethod
    def forward(ctx, density_grid, k0_grid, rays_o, rays_d, t_table, mask_world, cfg, ddesc, kdesc, dense_known, coherent=False):
        dev = rays_o.device
        rays_o = rays_o.contiguous().float()
        rays_d = rays_d.contiguous().float()
        N, S = rays_o.shape[0], cfg.n_samples
        f32 = dict(dtype=torch.float32, device=dev)
        dens = torch.empty(N * S, **f32)
        alpha = torch.empty(N * S, **f32)
        weight = torch.empty(N * S, **f32)
        T = torch.empty(N * S, **f32)
        flags = torch.empty(N * S, dtype=torch.uint8, device=dev)
        last = torch.empty(N, **f32)
        nkeep = torch.empty(N, dtype=torch.int32, device=dev)
        offsets = torch.empty(N + 1, dtype=torch.int64, device=dev)
        scratch = torch.empty(N // 1024 + 4, dtype=torch.int64, device=dev)
        with ops._Guard(rays_o) as lib:
            st = stream_of(rays_o)
            with _cabi.timed('march_density_fwd'):
                check(lib.ubn_march_density_fwd(ptr(rays_o), ptr(rays_d), ptr(t_table), ptr(density_grid), ddesc,
                                                ptr(mask_world), cfg, c_i64(N), ptr(dens), ptr(alpha), ptr(weight), ptr(T),
                                                ptr(flags), ptr(last), ptr(nkeep), st))
            check(lib.ubn_exclusive_scan_i32(ptr(nkeep), c_i64(N), ptr(offsets), ptr(scratch), st))
            # compacted size: known without a host sync when nothing can be masked out
            M = N * S if dense_known else int(offsets[N].item())
            C = k0_grid.shape[1]
            feat = torch.empty(M, C, **f32)
            o_dens = torch.empty(M, **f32)
            o_alpha = torch.empty(M, **f32)
            o_weight = torch.empty(M, **f32)
            ray_id = torch.empty(M, dtype=torch.int64, device=dev)
            step_id = torch.empty(M, dtype=torch.int64, device=dev)
            o_t = torch.empty(M, **f32)
            o_inner = torch.empty(M, dtype=torch.bool, device=dev)
            use_tma = coherent and tma_supported(k0_grid) and not torch.is_grad_enabled()
            with _cabi.timed('march_feature_fwd_tma' if use_tma else 'march_feature_fwd'):
                if use_tma:       # render path: bricks of the feature grid staged by TMA for 32 adjacent rays x 4 steps
                    check(lib.ubn_march_feature_fwd_tma(ptr(rays_o), ptr(rays_d), ptr(t_table), ptr(k0_grid), kdesc, cfg, c_i64(N),
                                                        ptr(flags), ptr(offsets), ptr(dens), ptr(alpha), ptr(weight), ptr(feat),
                                                        ptr(o_dens), ptr(o_alpha), ptr(o_weight), ptr(ray_id), ptr(step_id),
                                                        ptr(o_t), ptr(o_inner), ptr(TMA_STATS), st))
                else:
                    check(lib.ubn_march_feature_fwd(ptr(rays_o), ptr(rays_d), ptr(t_table), ptr(k0_grid), kdesc, cfg, c_i64(N),
                                                    ptr(flags), ptr(offsets), ptr(dens), ptr(alpha), ptr(weight), ptr(feat),
                                                    ptr(o_dens), ptr(o_alpha), ptr(o_weight), ptr(ray_id), ptr(step_id),
                                                    ptr(o_t), ptr(o_inner), st))
        ctx.save_for_backward(rays_o, rays_d, t_table, dens, alpha, weight, T, flags, last, offsets)
        ctx.cfg, ctx.ddesc, ctx.kdesc = cfg, ddesc, kdesc
        ctx.dmeta = (density_grid.shape, density_grid.stride())
        ctx.kmeta = (k0_grid.shape, k0_grid.stride())
        # persistent gradient buffers (dist.PeerTail / grid.attach_grad_buffer): the scatter adds straight into a buffer the
        # training loop owns (peer-mapped for the multi-GPU tail, zero at the start of a step) instead of a fresh zero-filled
        # 1.5 GB allocation per step; the parameter's .grad is pointed at it and autograd gets no gradient to accumulate
        ctx.dparam, ctx.kparam = density_grid, k0_grid
        ctx.mark_non_differentiable(ray_id, step_id, o_t, o_inner)
        return o_weight, last, o_alpha, o_dens, feat, ray_id, step_id, o_t, o_inner

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_weight, g_last, g_alpha, g_dens, g_feat, *unused):
        rays_o, rays_d, t_table, dens, alpha, weight, T, flags, last, offsets = ctx.saved_tensors
        dev = rays_o.device
        N = rays_o.shape[0]
        cont = lambda g: g.contiguous() if g is not None else None
        g_weight, g_last, g_alpha, g_dens, g_feat = map(cont, (g_weight, g_last, g_alpha, g_dens, g_feat))
        grad_d = grad_k = None
        with ops._Guard(rays_o) as lib:
            want_k = ctx.needs_input_grad[1] and g_feat is not None
            want_d = ctx.needs_input_grad[0]
            buf_d = getattr(ctx.dparam, '_ubn_grad_buffer', None) if want_d else None
            buf_k = getattr(ctx.kparam, '_ubn_grad_buffer', None) if want_k else None
            if want_d:
                grad_d = buf_d if buf_d is not None else torch.empty_strided(*ctx.dmeta, dtype=torch.float32, device=dev).zero_()
            if want_k:
                grad_k = buf_k if buf_k is not None else torch.empty_strided(*ctx.kmeta, dtype=torch.float32, device=dev).zero_()
            if want_k:
                with _cabi.timed('march_feature_bwd'):
                    check(lib.ubn_march_feature_bwd(ptr(rays_o), ptr(rays_d), ptr(t_table), ctx.kdesc, ctx.cfg, c_i64(N),
                                                    ptr(flags), ptr(offsets), ptr(g_feat), ptr(grad_k), stream_of(rays_o)))
            if want_d:
                with _cabi.timed('march_density_bwd'):
                    check(lib.ubn_march_density_bwd(ptr(rays_o), ptr(rays_d), ptr(t_table), ctx.ddesc, ctx.cfg, c_i64(N),
                                                    ptr(dens), ptr(alpha), ptr(weight), ptr(T), ptr(flags), ptr(last),
                                                    ptr(offsets), ptr(g_weight), ptr(g_alpha), ptr(g_dens), ptr(g_last),
                                                    ptr(grad_d), stream_of(rays_o)))
        if buf_d is not None:
            ctx.dparam.grad, grad_d = buf_d, None
        if buf_k is not None:
            ctx.kparam.grad, grad_k = buf_k, None
        return grad_d, grad_k, None, None, None, None, None, None, None, None, None
