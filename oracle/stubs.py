"""oracle/stubs.py -- TEST INFRASTRUCTURE ONLY.

Registers CPU stand-ins for the modules the reference's Python files import by bare name
(render_utils_cuda: grid.py:10, dvgo.py:13; total_variation_cuda: grid.py:11; adam_upd_cuda:
masked_adam.py:3; ub360_utils_cuda: dcvgo.py:15) plus the two un-vendored third-party packages on
the path (torch_scatter.segment_coo, torch_efficient_distloss.flatten_eff_distloss), all backed by
oracle/cpu_ref.py.  With these in ``sys.modules`` the reference's own ``FourierGrid/{grid,
FourierGrid_grid,dvgo,dcvgo,FourierGrid_model,masked_adam}.py`` import and run UNMODIFIED on CPU
from /root/reference -- that is how tests/golden/ is generated (oracle/make_golden.py).
/root/reference exists only in the build container, never on the GPU box.
"""
import sys
import types

from . import cpu_ref

_RENDER = ['infer_t_minmax', 'infer_n_samples', 'infer_ray_start_dir', 'sample_pts_on_rays',
           'sample_ndc_pts_on_rays', 'sample_bg_pts_on_rays', 'maskcache_lookup', 'raw2alpha',
           'raw2alpha_backward', 'raw2alpha_nonuni', 'raw2alpha_nonuni_backward', 'alpha2weight',
           'alpha2weight_backward']


def _mod(name, fns):
    m = types.ModuleType(name)
    m.__doc__ = 'CPU oracle stand-in (oracle/stubs.py)'
    for f in fns:
        setattr(m, f, getattr(cpu_ref, f))
    return m


def install(reference_root='/root/reference'):
    """Insert the stand-ins and put the reference checkout on sys.path. Returns the module dict."""
    mods = {
        'render_utils_cuda': _mod('render_utils_cuda', _RENDER),
        'total_variation_cuda': _mod('total_variation_cuda', ['total_variation_add_grad']),
        'adam_upd_cuda': _mod('adam_upd_cuda', ['adam_upd', 'masked_adam_upd', 'adam_upd_with_perlr']),
        'ub360_utils_cuda': _mod('ub360_utils_cuda', ['cumdist_thres']),
        'torch_scatter': _mod('torch_scatter', ['segment_coo', 'scatter_add']),
        'torch_efficient_distloss': _mod('torch_efficient_distloss', ['flatten_eff_distloss']),
    }
    for k, v in mods.items():
        sys.modules[k] = v
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return mods
