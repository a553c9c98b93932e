"""oracle/cpu_ref.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).

Python face of the CPU oracle for the FourierGrid / DVGO hot path of
sjtuytc/UnboundedNeRFPytorch @ 3d7008d:

* L0 ops (the reference's CUDA-only functions) -> plain-C restatement in ``oracle/ref_ops.c``
  (built by ``make -C oracle cpu`` into ``oracle/libubn_oracle.so``), exposed here with the exact
  module-function signatures of the reference extension (render_utils.cpp:170-184,
  total_variation.cpp:22-24, adam_upd.cpp:79-86, ub360_utils.cpp:20-22) on CPU tensors.
* L1/L2 semantics (DenseGrid / FourierGrid / MaskGrid / model forward) -> pure torch restatement on
  top of ``F.grid_sample`` -- the very ATen call the reference makes (grid.py:57,
  FourierGrid_grid.py:71,74) -- following grid.py:50-61, FourierGrid_grid.py:21-36,60-78,
  dcvgo.py:228-384, FourierGrid_model.py:509-672, dvgo.py:306-425.

Parity status: the reference has no tests or golden vectors.  This oracle is
pinned by running the reference's own Python files on top of it (oracle/make_golden.py ->
tests/golden/) and, on the GPU box, against the reference's own CUDA extension (oracle/_ref/).
"""
import ctypes
import functools
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libubn_oracle.so')
_lib = None


def build(force=False):
    """Compile oracle/ref_ops.c -> oracle/libubn_oracle.so with gcc (no GPU needed)."""
    src = os.path.join(_HERE, 'ref_ops.c')
    if force or (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-std=c11', '-ffp-contract=off', '-fno-fast-math', '-fPIC',
                               '-shared', '-o', _LIB_PATH, src, '-lm'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ubo_sample_pts_count.restype = ctypes.c_int64
        _lib.ubo_adam_step_size.restype = ctypes.c_float
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32(t):
    assert t.dtype == torch.float32 and t.device.type == 'cpu', 'oracle works on CPU fp32 tensors'
    return t.contiguous()


_i64 = ctypes.c_int64
_f = ctypes.c_float
_i = ctypes.c_int


# --------------------------------------------------------------------------------------
# render_utils_cuda surface (render_utils.cpp:170-184)
# --------------------------------------------------------------------------------------
def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    lib().ubo_infer_t_minmax(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _f(near), _f(far), _i64(n),
                             _p(t_min), _p(t_max))
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    rays_d, t_min, t_max = map(_f32, (rays_d, t_min, t_max))
    n = t_min.shape[0]
    out = torch.empty(n, dtype=torch.int64)
    lib().ubo_infer_n_samples(_p(rays_d), _p(t_min), _p(t_max), _f(stepdist), _i64(n), _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    rays_o, rays_d, t_min = map(_f32, (rays_o, rays_d, t_min))
    n = rays_o.shape[0]
    st, dr = torch.empty_like(rays_o), torch.empty_like(rays_o)
    lib().ubo_infer_ray_start_dir(_p(rays_o), _p(rays_d), _p(t_min), _i64(n), _p(st), _p(dr))
    return [st, dr]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min, t_max = torch.empty(n), torch.empty(n)
    n_steps = torch.empty(n, dtype=torch.int64)
    tot = lib().ubo_sample_pts_count(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _f(near), _f(far),
                                     _f(stepdist), _i64(n), _p(t_min), _p(t_max), _p(n_steps))
    pts = torch.empty(tot, 3)
    mask = torch.empty(tot, dtype=torch.bool)
    ray_id = torch.empty(tot, dtype=torch.int64)
    step_id = torch.empty(tot, dtype=torch.int64)
    lib().ubo_sample_pts_emit(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _p(t_min), _p(n_steps),
                              _f(stepdist), _i64(n), _p(pts), _p(mask), _p(ray_id), _p(step_id))
    return [pts, mask, ray_id, step_id, n_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    rays_o, rays_d, xyz_min, xyz_max = map(_f32, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, N_samples, 3)
    mask = torch.empty(n, N_samples, dtype=torch.bool)
    lib().ubo_sample_ndc_pts_on_rays(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), _i64(N_samples), _i64(n),
                                     _p(pts), _p(mask))
    return [pts, mask]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    rays_o, rays_d, t_max = map(_f32, (rays_o, rays_d, t_max))
    n = rays_o.shape[0]
    pts = torch.empty(n, N_samples, 3)
    lib().ubo_sample_bg_pts_on_rays(_p(rays_o), _p(rays_d), _p(t_max), _f(bg_preserve), _i64(N_samples), _i64(n),
                                    _p(pts))
    return pts


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    assert world.dtype == torch.bool
    world = world.contiguous()
    xyz, sc, sh = map(_f32, (xyz, xyz2ijk_scale, xyz2ijk_shift))
    n = xyz.shape[0]
    out = torch.zeros(n, dtype=torch.bool)
    if n:
        lib().ubo_maskcache_lookup(_p(world), _p(xyz), _p(out), _p(sc), _p(sh),
                                   _i64(world.shape[0]), _i64(world.shape[1]), _i64(world.shape[2]), _i64(n))
    return out


def raw2alpha(density, shift, interval):
    density = _f32(density)
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    lib().ubo_raw2alpha(_p(density), _f(float(shift)), _f(float(interval)), _p(None), _i64(density.numel()),
                        _p(exp_d), _p(alpha))
    return [exp_d, alpha]


def raw2alpha_nonuni(density, shift, interval):
    density, interval = _f32(density), _f32(interval)
    exp_d, alpha = torch.empty_like(density), torch.empty_like(density)
    lib().ubo_raw2alpha(_p(density), _f(float(shift)), _f(0.0), _p(interval), _i64(density.numel()),
                        _p(exp_d), _p(alpha))
    return [exp_d, alpha]


def raw2alpha_backward(exp_d, grad_back, interval):
    exp_d, grad_back = _f32(exp_d), _f32(grad_back)
    grad = torch.empty_like(exp_d)
    lib().ubo_raw2alpha_backward(_p(exp_d), _p(grad_back), _f(float(interval)), _p(None), _i64(exp_d.numel()), _p(grad))
    return grad


def raw2alpha_nonuni_backward(exp_d, grad_back, interval):
    exp_d, grad_back, interval = _f32(exp_d), _f32(grad_back), _f32(interval)
    grad = torch.empty_like(exp_d)
    lib().ubo_raw2alpha_backward(_p(exp_d), _p(grad_back), _f(0.0), _p(interval), _i64(exp_d.numel()), _p(grad))
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    alpha = _f32(alpha)
    ray_id = ray_id.contiguous()
    assert ray_id.dtype == torch.int64
    n = alpha.numel()
    weight, T = torch.empty_like(alpha), torch.empty_like(alpha)
    last = torch.empty(n_rays)
    i_start = torch.empty(n_rays, dtype=torch.int64)
    i_end = torch.empty(n_rays, dtype=torch.int64)
    lib().ubo_alpha2weight(_p(alpha), _p(ray_id), _i64(n), _i64(n_rays), _p(weight), _p(T), _p(last),
                           _p(i_start), _p(i_end))
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights, grad_last):
    alpha, weight, T, alphainv_last, grad_weights, grad_last = map(
        _f32, (alpha, weight, T, alphainv_last, grad_weights, grad_last))
    grad = torch.empty_like(alpha)
    lib().ubo_alpha2weight_backward(_p(alpha), _p(weight), _p(T), _p(alphainv_last), _p(i_start.contiguous()),
                                    _p(i_end.contiguous()), _i64(alpha.numel()), _i64(n_rays),
                                    _p(grad_weights), _p(grad_last), _p(grad))
    return grad


# --------------------------------------------------------------------------------------
# total_variation_cuda / adam_upd_cuda / ub360_utils_cuda surfaces
# --------------------------------------------------------------------------------------
def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode):
    assert param.is_contiguous() and grad.is_contiguous() and param.dim() == 5
    lib().ubo_total_variation_add_grad(_p(param), _p(grad), _f(float(wx)), _f(float(wy)), _f(float(wz)),
                                       _i64(param.shape[2]), _i64(param.shape[3]), _i64(param.shape[4]),
                                       _i64(param.numel()), _i(int(bool(dense_mode))))


def _adam(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, mode):
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_contiguous() and t.dtype == torch.float32
    lib().ubo_adam_upd(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(perlr), _i64(param.numel()),
                       _i(int(step)), _f(beta1), _f(beta2), _f(lr), _f(eps), _i(mode))


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 0)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 1)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, _f32(perlr), step, beta1, beta2, lr, eps, 2)


def cumdist_thres(dist, thres):
    dist = _f32(dist)
    mask = torch.zeros(dist.shape, dtype=torch.bool)
    lib().ubo_cumdist_thres(_p(dist), _f(float(thres)), _i64(dist.shape[0]), _i64(dist.shape[1]), _p(mask))
    return mask


# --------------------------------------------------------------------------------------
# third-party shims on the reference's path
# --------------------------------------------------------------------------------------
def segment_coo(src, index, out, reduce='sum'):
    """torch_scatter.segment_coo(reduce='sum') semantics: sorted-index segmented sum into ``out``
    (call sites dvgo.py:401,418; dcvgo.py:345,354,377; FourierGrid_model.py:640,666)."""
    assert reduce == 'sum'
    return out.index_add_(0, index, src)


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_add stand-in (imported, never called on the live path: dmpigo.py:11)."""
    if out is None:
        size = list(src.shape)
        size[dim] = int(dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0))
        out = torch.zeros(size, dtype=src.dtype)
    return out.index_add_(dim, index, src)


def flatten_eff_distloss(w, s, interval, ray_id):
    """torch_efficient_distloss.flatten_eff_distloss restated from the maths the reference keeps
    in-tree (dcvgo.py:387-409): (1/3)*interval*w^2 + 2*w*(s*w_prefix - ws_prefix), summed, / n_rays."""
    n_rays = int(ray_id.max()) + 1 if ray_id.numel() else 1
    w_prefix = torch.zeros_like(w)
    ws_prefix = torch.zeros_like(w)
    ws = w * s
    wc, wsc = torch.cumsum(w, 0), torch.cumsum(ws, 0)
    # exclusive prefix within each segment
    first = torch.ones_like(ray_id, dtype=torch.bool)
    first[1:] = ray_id[1:] != ray_id[:-1]
    seg_start_idx = torch.nonzero(first).flatten()
    seg_of = torch.cumsum(first.long(), 0) - 1
    base_w = (wc - w)[seg_start_idx][seg_of]
    base_ws = (wsc - ws)[seg_start_idx][seg_of]
    w_prefix = wc - w - base_w
    ws_prefix = wsc - ws - base_ws
    loss_uni = (1 / 3) * interval * w.pow(2)
    loss_bi = 2 * w * (s * w_prefix - ws_prefix)
    return (loss_bi.sum() + loss_uni.sum()) / n_rays


# --------------------------------------------------------------------------------------
# L1: grids (pure torch on F.grid_sample, the reference's own ATen call)
# --------------------------------------------------------------------------------------
def dense_grid_forward(grid, xyz, xyz_min, xyz_max):
    """grid.py:50-61.  grid [1,C,X,Y,Z]; xyz [...,3] -> [...,C] (squeezed if C==1)."""
    C = grid.shape[1]
    shape = xyz.shape[:-1]
    xyz = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    out = F.grid_sample(grid, ind_norm, mode='bilinear', align_corners=True)
    out = out.reshape(C, -1).T.reshape(*shape, C)
    if C == 1:
        out = out.squeeze(-1)
    return out


def nerf_pos_embed(x, num_freqs):
    """FourierGrid_grid.py:21-36 (logscale=True)."""
    freq_bands = 2 ** torch.linspace(0, num_freqs - 1, num_freqs)
    out = [x]
    for freq in freq_bands:
        out += [torch.sin(freq * x), torch.cos(freq * x)]
    return torch.cat(out, -1)


def fourier_grid_forward(grid, xyz, xyz_min, xyz_max, num_freqs):
    """FourierGrid_grid.py:60-78.  grid [1+2F,C,X,Y,Z] (num_freqs=F>0) or [1,C,X,Y,Z] (num_freqs<=0)."""
    C = grid.shape[1]
    shape = xyz.shape[:-1]
    xyz = xyz.reshape(1, 1, 1, -1, 3)
    ind_norm = ((xyz - xyz_min) / (xyz_max - xyz_min)).flip((-1,)) * 2 - 1
    if num_freqs > 0:
        pos = nerf_pos_embed(ind_norm, num_freqs)                      # [1,1,1,M,3*(1+2F)]
        P = 1 + 2 * num_freqs
        batch = pos.reshape(1, 1, 1, -1, P, 3).permute(4, 0, 1, 2, 3, 5).reshape(P, 1, 1, -1, 3)
        out = F.grid_sample(grid, batch, mode='bilinear', align_corners=True)
        out = out.mean(0).reshape(C, -1).T.reshape(*shape, C)
    else:
        out = F.grid_sample(grid, ind_norm, mode='bilinear', align_corners=True)
        out = out.reshape(C, -1).T.reshape(*shape, C)
    if C == 1:
        out = out.squeeze(-1)
    return out


# --------------------------------------------------------------------------------------
# L2: ray schedules and model forwards
# --------------------------------------------------------------------------------------
def contracted_t_schedule(world_len, stepsize, bg_len, t_boundary):
    """dcvgo.py:241-248 (t_boundary=2) / FourierGrid_model.py:524-532 (t_boundary=1.5)."""
    N_inner = int(2 / (2 + 2 * bg_len) * world_len / stepsize) + 1
    N_outer = N_inner
    b_inner = torch.linspace(0, t_boundary, N_inner + 1)
    b_outer = t_boundary / torch.linspace(1, 1 / 128, N_outer + 1)
    t = torch.cat([(b_inner[1:] + b_inner[:-1]) * 0.5, (b_outer[1:] + b_outer[:-1]) * 0.5])
    return t


def contracted_sample_ray(rays_o, rays_d, scene_center, scene_radius, t, bg_len, contracted_norm='inf'):
    """dcvgo.py:239-262 and FourierGrid_model.py:522-552 (same arithmetic; A = bg_len)."""
    ro = (rays_o - scene_center) / scene_radius
    rd = rays_d / rays_d.norm(dim=-1, keepdim=True)
    ray_pts = ro[:, None, :] + rd[:, None, :] * t[None, :, None]
    if contracted_norm == 'inf':
        norm = ray_pts.abs().amax(dim=-1, keepdim=True)
    elif contracted_norm == 'l2':
        norm = ray_pts.norm(dim=-1, keepdim=True)
    else:
        raise NotImplementedError
    inner_mask = (norm <= 1)
    ray_pts = torch.where(inner_mask, ray_pts, ray_pts / norm * ((1 + bg_len) - bg_len / norm))
    return ray_pts, inner_mask.squeeze(-1)


def view_embedding(viewdirs, viewfreq):
    """dcvgo.py:337-338 / FourierGrid_model.py:632-633."""
    emb = (viewdirs.unsqueeze(-1) * viewfreq).flatten(-2)
    return torch.cat([viewdirs, emb.sin(), emb.cos()], -1)


def rgbnet_forward(feat, w):
    """3-layer rgbnet (FourierGrid_model.py:234-241): Linear-ReLU-Linear-ReLU-Linear, weights dict
    with keys W1 [128,39], b1, W2 [128,128], b2, W3 [3,128], b3."""
    h = torch.relu(F.linear(feat, w['W1'], w['b1']))
    h = torch.relu(F.linear(h, w['W2'], w['b2']))
    return F.linear(h, w['W3'], w['b3'])


# --------------------------------------------------------------------------------------
# rays of a view (dvgo.py:492-557), torch CPU restatement
# --------------------------------------------------------------------------------------
def get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center', jitter=None):
    """dvgo.get_rays (dvgo.py:492-520) + viewdirs + ndc_rays(near=1, focal=K[0][0]) (dvgo.py:532-557).
    mode 'random' takes the uniform offsets as ``jitter`` [2,H,W] (plane 0 added to i, plane 1 to j) so it is testable."""
    K = torch.as_tensor(K, dtype=torch.float64)
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='ij')
    i, j = i.t().float(), j.t().float()
    if mode == 'center':
        i, j = i + 0.5, j + 0.5
    elif mode == 'random':
        i, j = i + jitter[0], j + jitter[1]
    elif mode != 'lefttop':
        raise NotImplementedError
    if flip_x:
        i = i.flip((1,))
    if flip_y:
        j = j.flip((0,))
    fx, fy, cx, cy = float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2])
    if inverse_y:
        dirs = torch.stack([(i - cx) / fx, (j - cy) / fy, torch.ones_like(i)], -1)
    else:
        dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    viewdirs = rays_d / rays_d.norm(dim=-1, keepdim=True)
    if ndc:
        near, focal = 1., fx
        t = -(near + rays_o[..., 2]) / rays_d[..., 2]
        rays_o = rays_o + t[..., None] * rays_d
        o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
        o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
        o2 = 1. + 2. * near / rays_o[..., 2]
        d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
        d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
        d2 = -2. * near / rays_o[..., 2]
        rays_o, rays_d = torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)
    return rays_o.contiguous(), rays_d.contiguous(), viewdirs.contiguous()


# --------------------------------------------------------------------------------------
# autograd wrappers over the C oracle (dvgo.py:430-488 semantics) and full model forwards
# --------------------------------------------------------------------------------------
def _autograd_fns(ext):
    """Raw2Alpha / Alphas2Weights autograd Functions (dvgo.py:430-488) over `ext`: this module (the C restatement, CPU) or a
    namespace holding the reference's own CUDA extension functions from oracle/_ref (GPU leg of the oracle)."""
    if id(ext) in _fn_cache:
        return _fn_cache[id(ext)][1]

    class _Raw2Alpha(torch.autograd.Function):
        @staticmethod
        def forward(ctx, density, shift, interval):
            exp_d, alpha = ext.raw2alpha(density, shift, interval)
            ctx.save_for_backward(exp_d)
            ctx.interval = float(interval)
            return alpha

        @staticmethod
        def backward(ctx, g):
            return ext.raw2alpha_backward(ctx.saved_tensors[0], g.contiguous(), ctx.interval), None, None

    class _Alphas2Weights(torch.autograd.Function):
        @staticmethod
        def forward(ctx, alpha, ray_id, n_rays):
            w, T, last, i_s, i_e = ext.alpha2weight(alpha, ray_id, n_rays)
            ctx.save_for_backward(alpha, w, T, last, i_s, i_e)
            ctx.n_rays = n_rays
            return w, last

        @staticmethod
        def backward(ctx, gw, gl):
            alpha, w, T, last, i_s, i_e = ctx.saved_tensors
            return ext.alpha2weight_backward(alpha, w, T, last, i_s, i_e, ctx.n_rays, gw.contiguous(), gl.contiguous()), None, None

    _fn_cache[id(ext)] = (ext, (_Raw2Alpha, _Alphas2Weights))      # keeps ext alive so the id stays unique
    return _fn_cache[id(ext)][1]


_fn_cache = {}


def model_forward(flavor, p, rays_o, rays_d, viewdirs, stepsize, bg=1, rand_bkgd=False, is_train=False,
                  render_depth=True, ext=None, keep_intermediates=False):
    """CPU restatement of FourierGridModel.forward (flavor='fouriergrid', FourierGrid_model.py:554-672) and
    DirectContractedVoxGO.forward (flavor='dcvgo', dcvgo.py:264-384).

    p: dict(density_grid [Pd,1,X,Y,Z], k0_grid [Pk,C,X,Y,Z], rgbnet=dict(W1,b1,W2,b2,W3,b3)|None, act_shift,
            scene_center[3], scene_radius[3], bg_len, contracted_norm, fast_color_thres, voxel_size_ratio,
            world_len, freq_density, freq_k0, viewfreq, mask (bool [X,Y,Z], dcvgo), mask_scale, mask_shift)
    Tensors with requires_grad=True in p receive gradients through the returned dict.

    Device-agnostic: with CPU tensors and ext=None the CUDA-only ops come from the C restatement in this module; with CUDA
    tensors and ext = the reference's own extension functions (oracle/_ref) this is the reference's GPU path op for op
    (ATen grid_sample, cuBLAS rgbnet, index_add in place of torch_scatter)."""
    import sys
    ext = ext if ext is not None else sys.modules[__name__]
    _Raw2Alpha, _Alphas2Weights = _autograd_fns(ext)
    dev = rays_o.device
    N = rays_o.shape[0]
    bg_len = p['bg_len']
    gmin = torch.tensor([-1., -1., -1.], device=dev) - bg_len
    gmax = torch.tensor([1., 1., 1.], device=dev) + bg_len
    t_boundary = 1.5 if flavor == 'fouriergrid' else 2.0
    t = contracted_t_schedule(p['world_len'], stepsize, bg_len, t_boundary).to(dev)
    ray_pts, inner_mask = contracted_sample_ray(rays_o, rays_d, p['scene_center'], p['scene_radius'], t, bg_len,
                                                p['contracted_norm'])
    S = len(t)
    interval = stepsize * float(p['voxel_size_ratio'])
    ray_id = torch.arange(N, device=dev).view(-1, 1).expand(N, S).flatten()
    step_id = torch.arange(S, device=dev).view(1, -1).expand(N, S).flatten()
    tt = t[None].repeat(N, 1)
    thres = p['fast_color_thres']

    def dgrid(x):
        return fourier_grid_forward(p['density_grid'], x, gmin, gmax, p['freq_density'])

    def kgrid(x):
        return fourier_grid_forward(p['k0_grid'], x, gmin, gmax, p['freq_k0'])

    if flavor == 'dcvgo':
        mask = inner_mask.clone()
        dist_thres = (2 + 2 * bg_len) / p['world_len'] * stepsize * 0.95
        dist = (ray_pts[:, 1:] - ray_pts[:, :-1]).norm(dim=-1)
        mask[:, 1:] |= ext.cumdist_thres(dist.contiguous(), dist_thres)
        ray_pts, inner_mask, tt = ray_pts[mask], inner_mask[mask], tt[mask]
        ray_id, step_id = ray_id[mask.flatten()], step_id[mask.flatten()]
        mask = ext.maskcache_lookup(p['mask'], ray_pts.contiguous(), p['mask_scale'], p['mask_shift'])
        ray_pts, inner_mask, tt, ray_id, step_id = ray_pts[mask], inner_mask[mask], tt[mask], ray_id[mask], step_id[mask]
    else:
        ray_pts, inner_mask, tt = ray_pts.reshape(-1, 3), inner_mask.reshape(-1), tt.reshape(-1)
    density = dgrid(ray_pts)
    density_q, pts_q = density, ray_pts          # every queried sample, before the threshold compactions (kept for the fp64 yardsticks)
    alpha = _Raw2Alpha.apply(density.contiguous(), float(p['act_shift']), interval)
    if thres > 0:
        mask = alpha > thres
        ray_pts, inner_mask, tt, ray_id, step_id = ray_pts[mask], inner_mask[mask], tt[mask], ray_id[mask], step_id[mask]
        density, alpha = density[mask], alpha[mask]
    weights, alphainv_last = _Alphas2Weights.apply(alpha.contiguous(), ray_id.contiguous(), N)
    if thres > 0:
        mask = weights > thres
        ray_pts, inner_mask, tt, ray_id, step_id = ray_pts[mask], inner_mask[mask], tt[mask], ray_id[mask], step_id[mask]
        density, alpha, weights = density[mask], alpha[mask], weights[mask]
    k0 = kgrid(ray_pts)
    if p.get('rgbnet') is None:
        rgb = torch.sigmoid(k0)
    else:
        emb = view_embedding(viewdirs, p['viewfreq']).flatten(0, -2)[ray_id]
        rgb = torch.sigmoid(rgbnet_forward(torch.cat([k0, emb], -1), p['rgbnet']))
    rgb_marched = torch.zeros(N, 3, device=dev).index_add_(0, ray_id, weights.unsqueeze(-1) * rgb)
    if flavor == 'dcvgo':
        if rand_bkgd and is_train:
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_marched)
        else:
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * bg
    elif rand_bkgd:
        rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_marched)
    s = 1 - 1 / (1 + tt)
    ret = dict(alphainv_last=alphainv_last, weights=weights, rgb_marched=rgb_marched, raw_density=density, raw_alpha=alpha,
               raw_rgb=rgb, ray_id=ray_id, step_id=step_id, n_max=S, t=tt, s=s)
    if flavor == 'dcvgo':
        ret['wsum_mid'] = torch.zeros(N, device=dev).index_add_(0, ray_id[inner_mask], weights[inner_mask])
    if render_depth:
        with torch.no_grad():
            ret['depth'] = torch.zeros(N, device=dev).index_add_(0, ray_id, weights * s)
    if keep_intermediates:      # for the fp64 re-evaluation of the colour branch in tests/parity_at_size.py
        ret['_ray_pts'], ret['_k0'] = ray_pts.detach(), k0.detach()
        ret['_density_q'], ret['_pts_q'] = density_q, pts_q.detach()     # graph tensor: d loss / d raw density of all queried samples
    return ret


def params_from_state(flavor, kwargs, state, requires_grad=False):
    """Build the ``p`` dict of model_forward from a reference-style (kwargs, state_dict) pair."""
    g = lambda k: state[k].detach().clone().float().requires_grad_(requires_grad)
    bg_len = kwargs.get('bg_len', 0.2)
    gmin = torch.tensor([-1., -1., -1.]) - bg_len
    gmax = torch.tensor([1., 1., 1.]) + bg_len
    p = dict(density_grid=g('density.grid'), k0_grid=g('k0.grid'), act_shift=float(state['act_shift']),
             scene_center=state['scene_center'].float(), scene_radius=state['scene_radius'].float(), bg_len=bg_len,
             contracted_norm=kwargs.get('contracted_norm', 'inf'), fast_color_thres=kwargs.get('fast_color_thres', 0))
    if 'rgbnet.0.weight' in state:
        p['rgbnet'] = dict(W1=g('rgbnet.0.weight'), b1=g('rgbnet.0.bias'), W2=g('rgbnet.2.0.weight'),
                           b2=g('rgbnet.2.0.bias'), W3=g('rgbnet.3.weight'), b3=g('rgbnet.3.bias'))
        p['viewfreq'] = state['viewfreq'].float()
    else:
        p['rgbnet'] = None
    if flavor == 'fouriergrid':
        nv, nvb = kwargs['num_voxels_density'], kwargs['num_voxels_base_density']
        F_ = kwargs.get('fourier_freq_num', 5)
        p['freq_density'] = F_
        p['freq_k0'] = F_ if p['rgbnet'] is not None else 0
    else:
        nv, nvb = kwargs['num_voxels'], kwargs['num_voxels_base']
        p['freq_density'] = p['freq_k0'] = 0
        p['mask'] = state['mask_cache.mask'].bool()
        p['mask_scale'] = state['mask_cache.xyz2ijk_scale'].float()
        p['mask_shift'] = state['mask_cache.xyz2ijk_shift'].float()
    voxel_size = ((gmax - gmin).prod() / nv).pow(1 / 3)
    voxel_size_base = ((gmax - gmin).prod() / nvb).pow(1 / 3)
    p['world_len'] = ((gmax - gmin) / voxel_size).long()[0].item()
    p['voxel_size_ratio'] = voxel_size / voxel_size_base
    return p
