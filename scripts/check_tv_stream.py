"""A/B of the streaming TV kernel against the element-per-thread one (UBN_TV_IMPL=0) and the reference's own CUDA build:
bit-level comparison + timing on the truck k0 grid.  Usage: python scripts/check_tv_stream.py  (GPU box)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(shape, dense, out):
    from unboundednerfpytorch_b200 import grid as G, ops
    g = torch.Generator().manual_seed(sum(shape))
    param = torch.randn(shape, generator=g) * 2
    grad = torch.randn(shape, generator=g) * (torch.rand(shape, generator=g) > 0.6)
    p, gr = G._as_cl3d(param.cuda()), G._as_cl3d(grad.cuda())
    ops.total_variation_add_grad(p, gr, 0.3, 0.2, 0.1, dense)
    res = gr.contiguous().cpu()
    if out:
        torch.save(res, out)
    return param, grad, res


def main():
    if len(sys.argv) > 1:                       # child: dump results of every case
        for n, (shape, dense) in enumerate(CASES):
            run_one(shape, dense, f'{sys.argv[1]}_{n}.pt')
        return
    from tests.util import ref_cuda
    tmp = '/tmp/tv_ab'
    for impl in ('0', '1'):
        subprocess.check_call([sys.executable, __file__, f'{tmp}{impl}'], env=dict(os.environ, UBN_TV_IMPL=impl))
    ref = ref_cuda('total_variation_cuda')
    for n, (shape, dense) in enumerate(CASES):
        a, b = torch.load(f'{tmp}0_{n}.pt'), torch.load(f'{tmp}1_{n}.pt')
        line = f'{shape} dense={dense}: stream==elementwise {torch.equal(a, b)} (max abs diff {(a - b).abs().max().item():.3e})'
        if ref is not None:
            g = torch.Generator().manual_seed(sum(shape))
            param = torch.randn(shape, generator=g) * 2
            grad = torch.randn(shape, generator=g) * (torch.rand(shape, generator=g) > 0.6)
            pc, gc = param.cuda().contiguous(), grad.cuda().contiguous()
            ref.total_variation_add_grad(pc, gc, 0.3, 0.2, 0.1, dense)
            line += f'; stream==reference-cuda {torch.equal(gc.cpu(), b)} (max abs diff {(gc.cpu() - b).abs().max().item():.3e})'
        print(line)
    # timing on the truck k0 grid
    from unboundednerfpytorch_b200 import grid as G, ops
    shape = (9, 12, 153, 153, 153)
    p = G._as_cl3d(torch.randn(shape, device='cuda'))
    gr = G._as_cl3d(torch.randn(shape, device='cuda'))
    for impl in ('stream (this process)',):
        for _ in range(3):
            ops.total_variation_add_grad(p, gr, 1e-3, 1e-3, 1e-3, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.total_variation_add_grad(p, gr, 1e-3, 1e-3, 1e-3, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'TV {impl} UBN_TV_IMPL={os.environ.get("UBN_TV_IMPL", "1")}: {ms:.3f} ms per sweep of {p.numel() / 1e6:.0f} M elements '
              f'({p.numel() * 12 / ms / 1e6:.0f} GB/s algorithmic at 12 B/elt)')


CASES = [((2, 12, 20, 9, 11), True), ((2, 12, 20, 9, 11), False), ((1, 12, 40, 70, 11), True), ((1, 12, 40, 70, 11), False),
         ((3, 4, 17, 33, 40), True), ((9, 12, 31, 30, 29), False)]

if __name__ == '__main__':
    main()
