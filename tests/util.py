"""Shared helpers for the parity tests."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5          # north-star tolerance for fp32 RGB / depth / weights (BASELINE.json)


def load_golden(name):
    return torch.load(os.path.join(ROOT, 'tests', 'golden', name), map_location='cpu', weights_only=False)


def assert_close(a, b, rtol=RTOL, atol=1e-6, what=''):
    if isinstance(rtol, str):              # allow assert_close(a, b, 'name') like assert_equal
        rtol, what = RTOL, rtol
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    if a.numel() == 0:
        return
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), (f'{what}: {int(bad.sum())}/{a.numel()} elements off; max abs err {err.max().item():.3e}, '
                           f'max rel err {(err / b.abs().clamp_min(1e-12)).max().item():.3e}')


def assert_equal(a, b, what=''):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    assert torch.equal(a, b), f'{what}: {int((a != b).sum())}/{a.numel()} elements differ (bit-exact target)'


_ref_cache = {}


def ref_cuda(name):
    """The reference's OWN CUDA extension module built into oracle/_ref (the GPU oracle: `make -C oracle ref`, compiled
    from /root/reference in the build container, shipped to the GPU box as a built .so).

    Never returns None on a CUDA box: a missing / unloadable oracle FAILS the calling test, so the "vs ref-cuda"
    assertions cannot be skipped silently (UBN_ALLOW_NO_REF=1 turns the failure into an explicit skip)."""
    if name in _ref_cache:
        return _ref_cache[name]
    path = os.path.join(ROOT, 'oracle', '_ref', f'{name}.so')
    why = None
    mod = None
    if not os.path.exists(path):
        why = f'{path} is missing (build it with `make -C oracle ref` where /root/reference exists)'
    else:
        try:
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except Exception as e:            # pragma: no cover
            why = f'could not load reference extension {path}: {e}'
    if mod is None:
        import pytest
        if os.environ.get('UBN_ALLOW_NO_REF') == '1':
            pytest.skip(f'reference CUDA oracle unavailable (UBN_ALLOW_NO_REF=1): {why}')
        pytest.fail(f'reference CUDA oracle unavailable -- the bit-exact "vs ref-cuda" checks would not run: {why}')
    print(f'[ref-cuda] loaded {path}')
    _ref_cache[name] = mod
    return mod


def ref_ext():
    """Namespace of the reference's own CUDA functions in the shape oracle.cpu_ref.model_forward(ext=...) expects: with CUDA
    tensors that call IS the reference's GPU path op for op (ATen grid_sample, cuBLAS rgbnet, index_add for torch_scatter,
    the reference's .cu kernels for everything else)."""
    import types
    ru, ub = ref_cuda('render_utils_cuda'), ref_cuda('ub360_utils_cuda')
    return types.SimpleNamespace(raw2alpha=ru.raw2alpha, raw2alpha_backward=ru.raw2alpha_backward, alpha2weight=ru.alpha2weight,
                                 alpha2weight_backward=ru.alpha2weight_backward, maskcache_lookup=ru.maskcache_lookup,
                                 cumdist_thres=ub.cumdist_thres)


def seeded_rays(n, seed, device='cpu', spread=0.5):
    g = torch.Generator().manual_seed(seed)
    ro = (torch.rand(n, 3, generator=g) - 0.5) * 2 * spread
    rd = torch.randn(n, 3, generator=g)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    return ro.to(device), rd.to(device), vd.to(device)
