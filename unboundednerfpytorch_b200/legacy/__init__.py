"""The four extension modules the reference imports by bare name (grid.py:10-11, dvgo.py:13, dcvgo.py:15,
masked_adam.py:3, FourierGrid_model.py:17-18).  ``install()`` makes ``import render_utils_cuda`` etc. resolve
to this package instead of the reference's ``python setup.py install`` build (README.md:138-144)."""
import importlib
import sys

NAMES = ('render_utils_cuda', 'total_variation_cuda', 'adam_upd_cuda', 'ub360_utils_cuda')


def install():
    mods = {}
    for n in NAMES:
        m = importlib.import_module(f'unboundednerfpytorch_b200.legacy.{n}')
        sys.modules[n] = m
        mods[n] = m
    return mods
