#!/usr/bin/env python
"""Diagnostic (GPU box): error statistics of the fused CUDA path against the reference's GPU path (oracle/_ref + ATen) on the
benchmarked configurations at 8192 x 512.  Prints one JSON object per configuration; asserts nothing (the asserting twin
is tests/test_gpu_parity_at_size.py).   python scripts/parity_at_size_report.py [config ...] > gpurun_out/parity.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import parity_at_size as P  # noqa: E402
from tests.util import ref_ext  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    ext = ref_ext()
    # torch.linspace CPU vs CUDA: the reference builds its t schedule on the default (CUDA) device, this library on the CPU
    from unboundednerfpytorch_b200 import march
    for wl, st, tb in ((153, 0.5, 1.5), (320, 1.045, 2.0), (200, 0.5, 1.5)):
        n_inner = int(2 / 2.4 * wl / st) + 1
        a = torch.linspace(0, tb, n_inner + 1)
        b = torch.linspace(0, tb, n_inner + 1, device=dev).cpu()
        c = tb / torch.linspace(1, 1 / 128, n_inner + 1)
        d = (tb / torch.linspace(1, 1 / 128, n_inner + 1, device=dev)).cpu()
        print(json.dumps({'linspace_check': [wl, st, tb], 'inner_equal': bool(torch.equal(a, b)), 'outer_equal': bool(torch.equal(c, d)),
                          'inner_mism': int((a != b).sum()), 'outer_mism': int((c != d).sum())}), flush=True)
    for name in (sys.argv[1:] or list(P.CONFIGS)):
        try:
            out, ours, p = P.compare(name, dev, ext=ext)
        except Exception as e:      # keep going: one OOM / failure must not hide the other configurations
            out = {'config': name, 'error': repr(e)}
        print(json.dumps(out), flush=True)
        ours = p = None
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
