"""Shared machinery of the at-size parity tests (tests/test_gpu_parity_at_size.py) and of the diagnostic script
scripts/parity_at_size_report.py: run the BENCHMARKED configurations (bench.py's workloads, 8192 rays x 512 samples) through

  (a) this library's fused CUDA path (models.*.forward + loss + backward), and
  (b) the reference's GPU path op for op: oracle.cpu_ref.model_forward on CUDA tensors with ext = the reference's OWN CUDA
      extension compiled into oracle/_ref (ATen grid_sample, cuBLAS rgbnet, index_add for torch_scatter) -- i.e. what
      FourierGrid_model.py:554-672 / dcvgo.py:264-384 execute on a GPU,

on identical seeded grids, rays and targets, and reduce the differences to a small dict of statistics.

north_star tolerance: sample indices / hit masks bit-exact; fp32 rgb / depth / weights within 1e-5 relative.  "Relative" is
taken against the larger of |reference value| and the tensor's scale (max |reference|): a per-element relative error of a
quantity that passes through zero (raw_density, gradients) is not meaningful, the usual rtol + atol = rtol * scale form is.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTOL = 1e-5
FLOAT_KEYS = ('rgb_marched', 'alphainv_last', 'depth', 'weights', 'raw_alpha', 'raw_density', 'raw_rgb', 't', 's', 'wsum_mid')

# name -> (bench workload, density N(mean, std), fast_color_thres, mask-cache keep probability)
CONFIGS = {
    'truck_dense': dict(workload='truck', mean=0.0, std=1.0, thres=0.0),             # the headline bench line (rho = 1)
    'truck_realistic': dict(workload='truck', mean=6.0, std=4.0, thres=1e-4),        # SURVEY 8d realistic mode: rays terminate
    'bicycle_dense': dict(workload='bicycle', mean=0.0, std=1.0, thres=0.0, mask_keep=0.9),   # DCVGO 320^3 + cumdist + mask cache
}


def build_pair(name, dev, n_rays=8192, seed=777):
    """-> (ours: nn.Module on dev, p: oracle parameter dict on dev, (ro, rd, vd, target) on dev, stepsize, flavor)."""
    import bench
    from oracle import cpu_ref
    from unboundednerfpytorch_b200 import models
    c = CONFIGS[name]
    flavor, kw, stepsize = bench.workload_kwargs(c['workload'])
    kw = dict(kw, fast_color_thres=c['thres'])
    torch.manual_seed(seed)
    cls = models.FourierGridModel if flavor == 'fouriergrid' else models.DirectContractedVoxGO
    m = cls(**kw)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        m.density.grid.copy_(torch.randn(m.density.grid.shape, generator=g) * c['std'] + c['mean'])
        m.k0.grid.copy_(torch.randn(m.k0.grid.shape, generator=g))
        if flavor == 'dcvgo':
            m.mask_cache.mask.copy_(torch.rand(m.mask_cache.mask.shape, generator=g) < c.get('mask_keep', 1.0))
    state = {k: v.detach().clone().contiguous() for k, v in m.state_dict().items()}        # reference layout [P,C,X,Y,Z]
    p = cpu_ref.params_from_state(flavor, kw, state, requires_grad=False)
    for k, v in list(p.items()):
        if torch.is_tensor(v):
            p[k] = v.to(dev)
    p['rgbnet'] = {k: v.to(dev).requires_grad_(True) for k, v in p['rgbnet'].items()}
    for k in ('density_grid', 'k0_grid'):
        p[k] = p[k].requires_grad_(True)
    del state
    ours = m.to(dev)
    batch = [t.to(dev) for t in bench.synth_batch(n_rays, seed)]
    return ours, p, batch, stepsize, flavor


def _stat(a, b):
    """Error statistics of a (ours) against b (reference GPU path):
    rel_scale = max |a-b| / max|b|                       (rtol * scale criterion)
    rel_elem  = max |a-b| / |b| over elements with |b| >= 1 % of the scale     (per-element relative error away from zero)
    frac_gt   = fraction of elements with |a-b| > 1e-5 * max(|b|, 1 % scale)."""
    a, b = a.detach().float(), b.detach().float().reshape(a.shape)
    if a.numel() == 0:
        return dict(n=0, max_abs=0.0, scale=0.0, rel_scale=0.0, rel_elem=0.0, frac_gt=0.0)
    scale = b.abs().max().item()
    err = (a - b).abs()
    floor = 0.01 * scale if scale > 0 else 1.0
    big = b.abs() >= floor
    rel_elem = (err[big] / b.abs()[big]).max().item() if bool(big.any()) else 0.0
    frac = (err > RTOL * b.abs().clamp_min(floor)).float().mean().item()
    return dict(n=a.numel(), max_abs=err.max().item(), scale=scale, rel_scale=err.max().item() / (scale if scale > 0 else 1.0),
                rel_elem=rel_elem, frac_gt=frac)


def compare(name, dev, n_rays=8192, backward=True, ext=None):
    """Run both paths, return {'ids_equal', 'M', 'M_ref', key: stat..., 'grad <param>': stat...}."""
    import bench
    from oracle import cpu_ref
    ours, p, (ro, rd, vd, target), stepsize, flavor = build_pair(name, dev, n_rays)
    out = {'config': name, 'flavor': flavor}
    rk = dict(near=0., far=1e9, bg=1, rand_bkgd=False, stepsize=stepsize, render_depth=True)
    ref = cpu_ref.model_forward(flavor, p, ro, rd, vd, stepsize, bg=1, rand_bkgd=False, render_depth=True, ext=ext)
    ret = ours(ro, rd, vd, global_step=None, is_train=False, **rk)
    out['M'], out['M_ref'], out['n_max'] = int(ret['ray_id'].numel()), int(ref['ray_id'].numel()), int(ret['n_max'])
    same_shape = ret['ray_id'].shape == ref['ray_id'].shape
    out['ray_id_equal'] = bool(same_shape and torch.equal(ret['ray_id'], ref['ray_id']))
    out['step_id_equal'] = bool(same_shape and torch.equal(ret['step_id'], ref['step_id']))
    if not same_shape or not (out['ray_id_equal'] and out['step_id_equal']):
        ka = ret['ray_id'] * 65536 + ret['step_id']
        kb = ref['ray_id'] * 65536 + ref['step_id']
        in_b, in_a = torch.isin(ka, kb), torch.isin(kb, ka)
        out['flips'] = int((~in_b).sum() + (~in_a).sum())
        for k in FLOAT_KEYS:                       # diagnostics on the common samples / all rays (the test fails on flips anyway)
            if k in ref and k in ret:
                if ret[k].shape[0] == ret['ray_id'].shape[0] and ret[k].dim() >= 1 and k not in ('rgb_marched', 'alphainv_last', 'depth', 'wsum_mid'):
                    out[k] = _stat(ret[k][in_b], ref[k].reshape(ref['ray_id'].shape[0], *ret[k].shape[1:])[in_a])
                else:
                    out[k] = _stat(ret[k], ref[k])
        return out, ours, p
    out['flips'] = 0
    for k in FLOAT_KEYS:
        if k in ref and k in ret:
            out[k] = _stat(ret[k], ref[k])
    if backward:
        ours.zero_grad(set_to_none=True)
        bench.step_loss(ret, target, n_rays).backward()
        bench.step_loss(ref, target, n_rays).backward()
        pairs = [('density.grid', ours.density.grid.grad, p['density_grid'].grad), ('k0.grid', ours.k0.grid.grad, p['k0_grid'].grad)]
        names = {'W1': ours.rgbnet[0].weight, 'b1': ours.rgbnet[0].bias, 'W2': ours.rgbnet[2][0].weight,
                 'b2': ours.rgbnet[2][0].bias, 'W3': ours.rgbnet[3].weight, 'b3': ours.rgbnet[3].bias}
        pairs += [('rgbnet.' + k, v.grad, p['rgbnet'][k].grad) for k, v in names.items()]
        for nm, a, b in pairs:
            out['grad ' + nm] = _stat(a, b)
    return out, ours, p
