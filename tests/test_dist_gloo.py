"""CPU, world_size 2 over gloo: the ray-sharding / gradient-exchange plumbing of the multi-GPU path (dist.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from unboundednerfpytorch_b200 import dist as D, grid as G
    r, w, _ = D.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    try:
        # 1) contiguous ray shards cover [0, n) exactly once
        n = 1001
        lo, hi = D.shard_range(n, rank, world)
        cover = torch.zeros(n)
        cover[lo:hi] = 1
        dist.all_reduce(cover)
        assert torch.equal(cover, torch.ones(n))
        rays = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
        (mine,) = D.shard_rays(rank, world, rays)
        assert torch.equal(mine, rays[lo:hi])
        # 2) frame gather: every rank renders its shard, all ranks get the assembled frame
        frame = D.gather_frame(mine * 2, n, rank, world)
        assert torch.equal(frame, rays * 2)
        # 3) gradient exchange: mean over ranks == single-process gradient of the mean loss on the concatenated batch, incl. a
        #    channels-last grid gradient (reduced in memory order without a copy) and a small rgbnet parameter
        torch.manual_seed(0)
        full = torch.randn(world, 2, 4, 3, 3, 3)
        p_grid = torch.nn.Parameter(G.zeros_grid([2, 4, 3, 3, 3]))
        p_grid.grad = G._as_cl3d(full[rank].clone())
        p_lin = torch.nn.Parameter(torch.zeros(5))
        p_lin.grad = torch.full((5,), float(rank + 1))
        p_none = torch.nn.Parameter(torch.zeros(2))           # no grad on this rank: skipped
        D.allreduce_grads([p_grid, p_lin, p_none])
        assert p_grid.grad.stride() == p_grid.stride()
        assert torch.allclose(p_grid.grad, full.mean(0))           # MEAN over ranks: per-rank losses are per-rank means
        assert torch.allclose(p_lin.grad, torch.full((5,), float(sum(range(1, world + 1))) / world))
        # 4) Block-NeRF style inverse-distance compositing of per-rank renders
        o = torch.zeros(7, 3)
        cent = torch.tensor([1.0 + rank, 0., 0.])
        rgb = torch.full((7, 3), float(rank))
        out = D.idw_composite(rgb, o, cent, power=4)
        wts = torch.tensor([(1.0 + k) ** -4 for k in range(world)])
        want = (wts * torch.arange(world)).sum() / wts.sum()
        assert torch.allclose(out, torch.full((7, 3), float(want)))
        q.put((rank, 'ok'))
    except Exception as e:                       # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _tail_worker(rank, world, port, q):
    """reduce_tv_step (slab-pipelined all-reduce -> TV -> Adam) == allreduce_grads -> TV -> opt.step(), bit for bit.
    The sweeps themselves are CUDA-only in the product, so on CPU the oracle's C restatement stands in for them."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle import cpu_ref
    from unboundednerfpytorch_b200 import dist as D, ops
    from unboundednerfpytorch_b200.masked_adam import MaskedAdam
    for name in ('total_variation_add_grad', 'adam_upd', 'masked_adam_upd', 'adam_upd_with_perlr'):
        setattr(ops, name, getattr(cpu_ref, name))
    D.init_from_env(backend='gloo')
    try:
        def make():
            torch.manual_seed(5)
            grid9 = torch.nn.Parameter(torch.randn(3, 2, 4, 5, 6))      # 3 slabs -> 3 pipelined chunks
            grid1 = torch.nn.Parameter(torch.randn(1, 2, 4, 4, 4))      # single slab -> reduced whole
            lin = torch.nn.Parameter(torch.randn(7, 3))
            opt = MaskedAdam([dict(params=[grid9, grid1], lr=0.1, skip_zero_grad=True),
                              dict(params=[lin], lr=1e-3, skip_zero_grad=False)])
            return (grid9, grid1, lin), opt
        (pa, oa), (pb, ob) = make(), make()
        for step in range(2):
            g = torch.Generator().manual_seed(100 * step + rank)
            for x, y in zip(pa, pb):
                grad = torch.randn(x.shape, generator=g) * (torch.rand(x.shape, generator=g) > 0.5)
                x.grad, y.grad = grad.clone(), grad.clone()
            tv_a = {pa[0]: (0.3, 0.2, 0.1, step == 0), pa[1]: (0.05, 0.05, 0.05, False)}
            D.reduce_tv_step(oa, tv_a)                                   # pipelined
            D.allreduce_grads(pb)                                        # sequential restatement
            ops.total_variation_add_grad(pb[0], pb[0].grad, 0.3, 0.2, 0.1, step == 0)
            ops.total_variation_add_grad(pb[1], pb[1].grad, 0.05, 0.05, 0.05, False)
            ob.step()
            for x, y in zip(pa, pb):
                assert torch.equal(x.grad, y.grad) and torch.equal(x.data, y.data)
                assert oa.state[x]['step'] == ob.state[y]['step'] == step + 1
                assert torch.equal(oa.state[x]['exp_avg'], ob.state[y]['exp_avg'])
                assert torch.equal(oa.state[x]['exp_avg_sq'], ob.state[y]['exp_avg_sq'])
        # replicas stay identical across ranks
        chk = pa[0].data.clone()
        dist.broadcast(chk, 0)
        assert torch.equal(chk, pa[0].data)
        q.put((rank, 'ok'))
    except Exception as e:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_pipelined_reduce_tv_adam_tail_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tail_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def _sparse_worker(rank, world, port, q):
    """allreduce_grads_sparse == allreduce_grads exactly; sparse route taken only where it pays."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from unboundednerfpytorch_b200 import dist as D, grid as G
    D.init_from_env(backend='gloo')
    try:
        g = torch.Generator().manual_seed(40 + rank)
        chunk = 64

        def make(shape, touched_frac, cl):
            # whole chunks of zeros IN MEMORY ORDER, different per rank; cl: logical [P,C,X,Y,Z] over a [P,X,Y,Z,C] buffer
            mem_shape = (shape[0], shape[2], shape[3], shape[4], shape[1]) if cl else shape
            full = torch.randn(mem_shape, generator=g)
            n = full.numel()
            keep = (torch.rand((n + chunk - 1) // chunk, generator=g) < touched_frac).repeat_interleave(chunk)[:n]
            full = (full.view(-1) * keep).view(mem_shape)
            return full.permute(0, 4, 1, 2, 3) if cl else full

        specs = [((3, 4, 9, 10, 11), 0.1, True),      # channels-last grid, sparse  -> sparse route (ragged tail: 11880 % 64 != 0)
                 ((1, 1, 20, 20, 20), 0.9, False),    # mostly touched             -> dense fallback
                 ((2, 2, 16, 16, 8), 0.0, False),     # nothing touched anywhere   -> nothing exchanged but flags
                 ((40,), 1.0, False)]                 # small                      -> dense
        pa, pb = [], []
        for shape, frac, cl in specs:
            grad = make(shape, frac, cl)
            for lst in (pa, pb):
                p = torch.nn.Parameter(torch.zeros(shape) if not cl else G.zeros_grid(list(shape)))
                p.grad = torch.empty_like(p, memory_format=torch.preserve_format).copy_(grad)
                assert p.grad.stride() == p.stride()
                lst.append(p)
        stats = D.allreduce_grads_sparse(pa, chunk=chunk, dense_above=0.5)
        D.allreduce_grads(pb)
        for x, y in zip(pa, pb):
            assert x.grad.stride() == y.grad.stride() and torch.equal(x.grad, y.grad)
        assert 0.0 < stats[pa[0]] <= 0.25 and stats[pa[1]] == 1.0 and stats[pa[2]] == 0.0 and pa[3] not in stats
        q.put((rank, 'ok'))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_chunk_sparse_grad_exchange_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_ray_sharding_and_grad_exchange_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_shard_range_properties():
    from unboundednerfpytorch_b200.dist import shard_range
    for n in (0, 1, 7, 8192, 1707200):
        for w in (1, 2, 4, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
