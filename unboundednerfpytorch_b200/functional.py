"""Autograd Functions with the reference's names and call signatures (FourierGrid/dvgo.py:430-488).

``Raw2Alpha.apply(density, shift, interval)``, ``Raw2Alpha_nonuni.apply(density, shift, interval)``,
``Alphas2Weights.apply(alpha, ray_id, N) -> (weights, alphainv_last)``.  ``shift`` / ``interval`` may be
Python numbers or 1-element tensors (the reference passes ``self.act_shift``, a 1-element CUDA tensor;
its pybind float caster turns that into an implicit ``.item()`` on EVERY call, SURVEY.md 8b -- here the
host value is cached per tensor version so steady-state calls do not synchronise).
"""
import weakref

import torch

from . import ops

_scalar_cache = {}


def host_scalar(v):
    """float(v); for device tensors the D2H read is cached per tensor OBJECT (weakref + in-place version counter), so a
    long-lived buffer such as `act_shift` costs one sync, and a new tensor that reuses a freed address never hits."""
    if isinstance(v, torch.Tensor) and v.is_cuda:
        hit = _scalar_cache.get(id(v))
        if hit is not None and hit[0]() is v and hit[1] == v._version:
            return hit[2]
        if len(_scalar_cache) > 64:
            _scalar_cache.clear()
        val = float(v)
        _scalar_cache[id(v)] = (weakref.ref(v), v._version, val)
        return val
    return float(v)


class Raw2Alpha(torch.autograd.Function):
    """alpha = 1 - (1 + exp(density + shift)) ** (-interval)   (dvgo.py:430-454)."""

    @staticmethod
    def forward(ctx, density, shift, interval):
        interval = host_scalar(interval)
        exp, alpha = ops.raw2alpha(density, host_scalar(shift), interval)
        if density.requires_grad:
            ctx.save_for_backward(exp)
            ctx.interval = interval
        return alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_back):
        exp = ctx.saved_tensors[0]
        return ops.raw2alpha_backward(exp, grad_back.contiguous(), ctx.interval), None, None


class Raw2Alpha_nonuni(torch.autograd.Function):
    """Per-point interval variant (dvgo.py:456-470)."""

    @staticmethod
    def forward(ctx, density, shift, interval):
        exp, alpha = ops.raw2alpha_nonuni(density, host_scalar(shift), interval)
        if density.requires_grad:
            ctx.save_for_backward(exp, interval)
        return alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_back):
        exp, interval = ctx.saved_tensors
        return ops.raw2alpha_nonuni_backward(exp, grad_back.contiguous(), interval), None, None


class Alphas2Weights(torch.autograd.Function):
    """weights_i = alpha_i * prod_{j<i}(1 - alpha_j) per ray with the reference's early stop (dvgo.py:472-488)."""

    @staticmethod
    def forward(ctx, alpha, ray_id, N):
        weights, T, alphainv_last, i_start, i_end = ops.alpha2weight(alpha, ray_id, N)
        if alpha.requires_grad:
            ctx.save_for_backward(alpha, weights, T, alphainv_last, i_start, i_end)
            ctx.n_rays = N
        return weights, alphainv_last

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_weights, grad_last):
        alpha, weights, T, alphainv_last, i_start, i_end = ctx.saved_tensors
        grad = ops.alpha2weight_backward(alpha, weights, T, alphainv_last, i_start, i_end, ctx.n_rays,
                                         grad_weights.contiguous(), grad_last.contiguous())
        return grad, None, None


class SegmentSum(torch.autograd.Function):
    """out[r] = sum of src rows whose (sorted) ray_id == r.  Forward: deterministic warp-per-ray reduction kernel
    (libubnerf_b200: ubn_segment_sum); backward: grad_src[i] = grad_out[ray_id[i]] (a gather)."""

    @staticmethod
    def forward(ctx, src, ray_id, n_rays):
        from . import _cabi
        from ._cabi import c_i64, check, ptr, stream_of
        if not src.is_cuda:
            raise RuntimeError('src must be a CUDA tensor')
        squeeze = src.dim() == 1
        s2 = (src.unsqueeze(-1) if squeeze else src).contiguous().float()
        k = s2.shape[1]
        ray_id = ray_id.contiguous()
        out = torch.empty(n_rays, k, dtype=torch.float32, device=src.device)
        i_s = torch.empty(n_rays, dtype=torch.int64, device=src.device)
        i_e = torch.empty(n_rays, dtype=torch.int64, device=src.device)
        with ops._Guard(src) as lib:
            check(lib.ubn_segment_sum(ptr(s2), c_i64(k), ptr(ray_id), c_i64(s2.shape[0]), c_i64(n_rays), ptr(i_s), ptr(i_e),
                                      ptr(out), stream_of(src)))
        ctx.save_for_backward(ray_id)
        ctx.squeeze = squeeze
        return out.squeeze(-1) if squeeze else out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (ray_id,) = ctx.saved_tensors
        return g[ray_id], None, None


def segment_sum(src, ray_id, n_rays):
    return SegmentSum.apply(src, ray_id, n_rays)


class CompositeRGB(torch.autograd.Function):
    """rgb_marched[r] = sum_{i: ray_id[i]==r} weights[i] * rgb[i]  ==  segment_sum(weights[:,None] * rgb, ray_id, n_rays)
    (FourierGrid_model.py:640-644) with the product never materialised; backward writes both gradients in one launch."""

    @staticmethod
    def forward(ctx, weights, rgb, ray_id, n_rays):
        from ._cabi import c_i64, check, ptr, stream_of
        if not (weights.is_cuda and rgb.is_cuda):
            raise RuntimeError('weights / rgb must be CUDA tensors')
        weights, rgb, ray_id = weights.contiguous().float(), rgb.contiguous().float(), ray_id.contiguous()
        assert rgb.dim() == 2 and rgb.shape[1] == 3 and weights.shape[0] == rgb.shape[0]
        out = torch.empty(n_rays, 3, dtype=torch.float32, device=rgb.device)
        i_s = torch.empty(n_rays, dtype=torch.int64, device=rgb.device)
        i_e = torch.empty(n_rays, dtype=torch.int64, device=rgb.device)
        with ops._Guard(rgb) as lib:
            check(lib.ubn_composite_fwd(ptr(weights), ptr(rgb), ptr(ray_id), c_i64(rgb.shape[0]), c_i64(n_rays), ptr(i_s), ptr(i_e),
                                        ptr(out), stream_of(rgb)))
        ctx.save_for_backward(weights, rgb, ray_id)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        from ._cabi import c_i64, check, ptr, stream_of
        weights, rgb, ray_id = ctx.saved_tensors
        g = g.contiguous()
        gw = torch.empty_like(weights) if ctx.needs_input_grad[0] else None
        grgb = torch.empty_like(rgb) if ctx.needs_input_grad[1] else None
        with ops._Guard(rgb) as lib:
            check(lib.ubn_composite_bwd(ptr(weights), ptr(rgb), ptr(ray_id), ptr(g), c_i64(rgb.shape[0]), ptr(gw), ptr(grgb),
                                        stream_of(rgb)))
        return gw, grgb, None, None


def composite_rgb(weights, rgb, ray_id, n_rays):
    return CompositeRGB.apply(weights, rgb, ray_id, n_rays)


def segment_coo(src, index, out, reduce='sum'):
    """torch_scatter.segment_coo(reduce='sum') for sorted ``index`` (call sites dvgo.py:401,418;
    dcvgo.py:345,354,377; FourierGrid_model.py:640,666): out += per-segment sums of the rows of ``src``."""
    if reduce != 'sum':
        raise NotImplementedError(reduce)
    return out + segment_sum(src, index, out.shape[0]).reshape(out.shape)


# --------------------------------------------------------------------------------------------------
# in-kernel training losses (run_train.py:254-279): MSE + entropy_last + per-point rgb loss
# --------------------------------------------------------------------------------------------------
class RenderLoss(torch.autograd.Function):
    """loss = w_main * mse(rgb_marched, target) + w_freq * FourierMSE(rgb_marched, target) + w_entropy * entropy_last(alphainv_last)
    + w_nearclip * nearclip(raw_density, t) + w_rgbper * rgbper, value and gradients from two launches (ubn_render_loss).
    Returns a [5] tensor {loss, mse, entropy_last, rgbper, freq}; only element 0 carries gradient (the others are the detached
    terms the training loop logs, e.g. psnr = mse2psnr(out[1]))."""

    @staticmethod
    def forward(ctx, rgb_marched, alphainv_last, raw_rgb, weights, ray_id, target, raw_density, t_pts, w_main, w_entropy, w_rgbper,
                w_freq, w_nearclip, near_thres):
        from ._cabi import c_f, c_i64, check, ptr, stream_of
        dev = rgb_marched.device
        for t, nm in ((rgb_marched, 'rgb_marched'), (target, 'target')):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f'{nm} must be a contiguous CUDA fp32 tensor')
        n_rays = rgb_marched.shape[0]
        use_ent = alphainv_last is not None and w_entropy != 0
        use_per = raw_rgb is not None and w_rgbper != 0 and raw_rgb.shape[0] > 0
        use_clip = raw_density is not None and t_pts is not None and w_nearclip != 0 and raw_density.numel() > 0
        if use_ent:
            alphainv_last = alphainv_last.contiguous()
        if use_per:
            raw_rgb, weights, ray_id = raw_rgb.contiguous(), weights.detach().contiguous(), ray_id.contiguous()
        if use_clip:
            t_pts = t_pts.contiguous()
        n_pts = raw_rgb.shape[0] if use_per else (raw_density.numel() if use_clip else 0)
        scratch = torch.empty(4 * 1024, dtype=torch.float64, device=dev)   # per call: safe under concurrent streams
        out = torch.empty(5, device=dev)
        g_rgb = torch.empty_like(rgb_marched)
        g_last = torch.empty_like(alphainv_last) if use_ent else None
        g_raw = torch.empty_like(raw_rgb) if use_per else None
        g_dens = torch.empty(raw_density.shape, dtype=torch.float32, device=dev) if use_clip else None
        with ops._Guard(rgb_marched) as lib:
            check(lib.ubn_render_loss(ptr(rgb_marched), ptr(alphainv_last if use_ent else None), ptr(raw_rgb if use_per else None),
                                      ptr(weights if use_per else None), ptr(ray_id if use_per else None), ptr(target),
                                      ptr(t_pts if use_clip else None), c_i64(n_rays), c_i64(n_pts), c_f(float(w_main)),
                                      c_f(float(w_entropy)), c_f(float(w_rgbper)), c_f(float(w_freq)),
                                      c_f(float(w_nearclip) if use_clip else 0.0), c_f(float(near_thres) if use_clip else 0.0),
                                      ptr(out), ptr(g_rgb), ptr(g_last), ptr(g_raw), ptr(g_dens), ptr(scratch),
                                      c_i64(scratch.numel()), stream_of(rgb_marched)))
        ctx.save_for_backward(g_rgb, g_last, g_raw, g_dens)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        g_rgb, g_last, g_raw, g_dens = ctx.saved_tensors
        s = grad_out[0]
        sc = lambda g: None if g is None else g * s
        return (sc(g_rgb), sc(g_last), sc(g_raw), None, None, None, sc(g_dens), None, None, None, None, None, None, None)


def render_loss(render_result, target, weight_main=1.0, weight_entropy_last=0.0, weight_rgbper=0.0, weight_freq=0.0,
                weight_nearclip=0.0, near_thres=None, weight_distortion=0.0):
    """The loss terms of the reference training loop (run_train.py:253-279) from a model ``ret_dict``:
    main (mse) + freq (FourierMSELoss) + entropy_last + nearclip + rgbper in two launches, + distortion (flatten_eff_distloss,
    one warp-per-ray kernel) when ``weight_distortion`` > 0.  ``near_thres`` = data_dict['near_clip'] / model.scene_radius[0]
    (run_train.py:263).  Returns (loss, {'mse', 'entropy_last', 'rgbper', 'freq', 'distortion'}) -- loss is differentiable, the
    terms are detached scalars."""
    clip = weight_nearclip > 0 and near_thres is not None
    out = RenderLoss.apply(render_result['rgb_marched'], render_result.get('alphainv_last'), render_result.get('raw_rgb'),
                           render_result.get('weights'), render_result.get('ray_id'), target.contiguous(),
                           render_result.get('raw_density') if clip else None, render_result.get('t') if clip else None,
                           weight_main, weight_entropy_last, weight_rgbper, weight_freq, weight_nearclip if clip else 0.0,
                           near_thres if clip else 0.0)
    d = out.detach()
    loss = out[0]
    terms = {'mse': d[1], 'entropy_last': d[2], 'rgbper': d[3], 'freq': d[4]}
    if weight_distortion > 0:
        dl = flatten_eff_distloss(render_result['weights'], render_result['s'], 1 / render_result['n_max'], render_result['ray_id'],
                                  n_rays=render_result['rgb_marched'].shape[0])
        loss = loss + weight_distortion * dl
        terms['distortion'] = dl.detach()
    return loss, terms


# --------------------------------------------------------------------------------------------------
# distortion loss (torch_efficient_distloss.flatten_eff_distloss, run_train.py:268-274)
# --------------------------------------------------------------------------------------------------
class DistortionLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, s, interval, ray_id, n_rays):
        from ._cabi import c_f, c_i64, check, ptr, stream_of
        if not (w.is_cuda and s.is_cuda and ray_id.is_cuda):
            raise RuntimeError('w / s / ray_id must be CUDA tensors')
        w, s, ray_id = w.contiguous().float(), s.contiguous().float(), ray_id.contiguous()
        dev = w.device
        out = torch.empty(1, device=dev)
        gw = torch.empty_like(w)
        i_s = torch.empty(n_rays, dtype=torch.int64, device=dev)
        i_e = torch.empty(n_rays, dtype=torch.int64, device=dev)
        scratch = torch.empty(max(n_rays, 1), dtype=torch.float64, device=dev)
        with ops._Guard(w) as lib:
            check(lib.ubn_distortion_loss(ptr(w), ptr(s), ptr(ray_id), c_i64(w.shape[0]), c_i64(n_rays), c_f(float(interval)),
                                          ptr(i_s), ptr(i_e), ptr(out), ptr(gw), ptr(scratch), c_i64(scratch.numel()),
                                          stream_of(w)))
        ctx.save_for_backward(gw)
        return out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (gw,) = ctx.saved_tensors
        return gw * g, None, None, None, None


def flatten_eff_distloss(w, s, interval, ray_id, n_rays=None):
    """Same call as torch_efficient_distloss.flatten_eff_distloss (run_train.py:273): mean over ray_id.max()+1 rays (the kernel
    reads that count on the device).  ``n_rays``: an upper bound of it (the batch size) used only to size the per-ray scratch;
    without it one host read of ray_id[-1] supplies the bound."""
    if w.numel() == 0:
        return w.sum() * 0
    if n_rays is None:
        n_rays = int(ray_id[-1]) + 1                  # sorted ids: the last one is the maximum
    return DistortionLoss.apply(w, s, interval, ray_id, n_rays)
