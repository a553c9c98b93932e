#!/usr/bin/env python
"""GPU probe: in which order does torch-CUDA evaluate `x.mean(0)` on the [P, C, 1, 1, M] output of F.grid_sample
(FourierGrid_grid.py:71-72)?  Prints mismatch counts of candidate fp32 evaluation orders against torch's own result."""
import json
import sys

import torch

dev = 'cuda:0'


def seq(x):
    s = x[0].clone()
    for i in range(1, x.shape[0]):
        s = s + x[i]
    return s


def acc4(x):
    P = x.shape[0]
    acc = [None] * 4
    for i in range(P):
        a = i % 4
        acc[a] = x[i].clone() if acc[a] is None else acc[a] + x[i]
    s = acc[0]
    for a in range(1, 4):
        if acc[a] is not None:
            s = s + acc[a]
    return s


def acc4_blocked(x):      # vt0 = 4 values per loop iteration, tail handled one by one into accumulator slots 0..
    P = x.shape[0]
    acc = [None] * 4
    i = 0
    while i + 4 <= P:
        for a in range(4):
            acc[a] = x[i + a].clone() if acc[a] is None else acc[a] + x[i + a]
        i += 4
    a = 0
    while i < P:
        acc[a] = x[i].clone() if acc[a] is None else acc[a] + x[i]
        i += 1
        a += 1
    s = acc[0]
    for a in range(1, 4):
        if acc[a] is not None:
            s = s + acc[a]
    return s


def tree(x):
    xs = [x[i] for i in range(x.shape[0])]
    while len(xs) > 1:
        nxt = [xs[i] + xs[i + 1] for i in range(0, len(xs) - 1, 2)]
        if len(xs) % 2:
            nxt.append(xs[-1])
        xs = nxt
    return xs[0]


for P in (3, 5, 7, 9, 11):
    for C, M in ((1, 1 << 20), (12, 1 << 18), (1, 4194304), (12, 4194304 // 4)):
        g = torch.Generator(device=dev).manual_seed(P * 100 + C)
        x = torch.randn(P, C, 1, 1, M, generator=g, device=dev)
        want = x.mean(0)
        inv = torch.tensor(1.0 / P, dtype=torch.float32, device=dev)        # fp32 reciprocal
        inv_d = 1.0 / P                                                    # python double, narrowed by torch when multiplied
        res = {}
        for name, fn in (('seq', seq), ('acc4', acc4), ('acc4_blocked', acc4_blocked), ('tree', tree)):
            s = fn(x)
            res[name + '/div'] = int((s / P != want).sum())
            res[name + '*inv32'] = int((s * inv != want).sum())
            res[name + '*invd'] = int((s * inv_d != want).sum())
        res['sum/div'] = int((x.sum(0) / P != want).sum())
        res['sum*inv32'] = int((x.sum(0) * inv != want).sum())
        print(json.dumps({'P': P, 'C': C, 'M': M, 'mismatches': res}), flush=True)
