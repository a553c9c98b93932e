"""Multi-GPU plumbing: one process per GPU, rays sharded, grids replicated.

The reference has no distributed code on this path (0 collectives in the tree).  Rays are
independent given the grids, so rendering needs no exchange until the frame is gathered, and training needs
exactly one: the sum of the grid / rgbnet gradients before TV + Adam, after which every rank applies the
identical update to its replica.  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / MASTER_*). Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            opts = None
            if os.environ.get('UBN_NCCL_HIGH_PRIORITY', '1') != '0':
                # collectives run on a high-priority stream so that their few CTAs are placed ahead of the queued CTAs of
                # the full-grid sweeps they overlap with (reduce_tv_step)
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group(backend=backend, device_id=torch.device('cuda', local), pg_options=opts)
        else:
            dist.init_process_group(backend=backend)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (contiguous keeps image-space coherence of a frame's rays);
    the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rank, world, *tensors):
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return tuple(t[lo:hi] for t in tensors)


class _MeanReduce:
    """All-reduce that leaves the MEAN over ranks in the tensor.

    Every loss term of the training step is a per-rank mean over that rank's rays (F.mse_loss, entropy_last .mean(), rgbper
    / len(rays_o), TV weight / len(rays_o): run_train.py:254-287), so the gradient of the same loss on the concatenated
    global batch of world x n rays is the mean -- not the sum -- of the per-rank gradients.  NCCL reduces with
    ReduceOp.AVG (free); gloo has no AVG, so the sum is scaled after the wait."""

    def __init__(self, tensor, async_op=True):
        self.t = tensor
        self.world = dist.get_world_size()
        self.native = dist.get_backend() == 'nccl'
        self.h = dist.all_reduce(tensor, op=dist.ReduceOp.AVG if self.native else dist.ReduceOp.SUM, async_op=async_op)

    def wait(self):
        if self.h is not None:
            self.h.wait()
            self.h = None
        if not self.native:
            self.t.mul_(1.0 / self.world)
            self.native = True


def _normalise_grad_layout(param):
    """Make ``param.grad`` share the parameter's strides (channels-last grids keep channels-last gradients) so that TV, the
    collectives and MaskedAdam all see one layout; autograd normally guarantees this, a torch-produced or accumulated
    gradient may not."""
    g = param.grad
    if g is not None and g.stride() != param.stride():
        param.grad = torch.empty_like(param, memory_format=torch.preserve_format).copy_(g)
    return param.grad


def allreduce_grads(params, world=None):
    """Average gradients across ranks in place (one collective per tensor, large grids first).  With identical replicas,
    ray-sharded batches of equal size and per-rank mean losses this reproduces the single-process gradient of the same
    loss on the concatenated batch up to fp32 reassociation (see _MeanReduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    handles = []
    for p in sorted((p for p in params if p.grad is not None), key=lambda p: -p.numel()):
        # channels-last grids: reduce the dense storage in memory order (no copy)
        handles.append(_MeanReduce(_memory_order(_normalise_grad_layout(p))))
    for h in handles:
        h.wait()


def allreduce_grads_sparse(params, chunk=4096, dense_above=0.5):
    """Chunk-sparse gradient exchange (SURVEY.md 8e "brick-sparse all-reduce", with 1-D bricks of the memory-order buffer so
    that no re-layout or padding of the grid is needed): for every large gradient
      1. per-chunk touch flags  (any element != 0 in a run of `chunk` consecutive floats)     -- one pass over the gradient
      2. union of the flags over ranks                                                        -- all-reduce(MAX) of n/chunk bytes
      3. gather the union's chunks, all-reduce only those, scatter them back                  -- |union| * chunk floats
    falling back to the dense all-reduce when the union covers more than `dense_above` of the tensor (then the
    compaction would cost more than it saves; the synthetic bench rays touch nearly every chunk) or the tensor is
    small.  Elements outside the union are zero on every rank, so the result equals the dense mean exactly, and
    MaskedAdam's "skip where the summed grad == 0" rule is unchanged.  One host read per large tensor (the union size).
    Returns {param: fraction of chunks exchanged} for the tensors that took the sparse route (1.0 = dense fallback)."""
    stats = {}
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    dense = []
    world = dist.get_world_size()
    for p in sorted((p for p in params if p.grad is not None), key=lambda p: -p.numel()):
        g = _memory_order(_normalise_grad_layout(p))
        n = g.numel()
        if n < 64 * chunk or not g.is_contiguous():
            dense.append(g)
            continue
        flat = g.view(-1)
        n_full = n // chunk
        body = flat[:n_full * chunk].view(n_full, chunk)
        touched = (body != 0).any(dim=1).to(torch.uint8)
        dist.all_reduce(touched, op=dist.ReduceOp.MAX)
        idx = touched.nonzero(as_tuple=False).squeeze(1)             # host sync: the size of the union
        frac = idx.numel() / max(n_full, 1)
        if frac > dense_above:
            dense.append(g)
            stats[p] = 1.0
            continue
        stats[p] = frac
        if idx.numel():
            buf = body.index_select(0, idx)
            _MeanReduce(buf, async_op=False).wait()
            body.index_copy_(0, idx, buf)
        if n_full * chunk < n:                                       # ragged tail: always exchanged
            _MeanReduce(flat[n_full * chunk:], async_op=False).wait()
    handles = [_MeanReduce(g) for g in dense]
    for h in handles:
        h.wait()
    return stats


def _memory_order(t):
    """Contiguous view of a channels-last 5-D grid (or the tensor itself) for a collective."""
    if not t.is_contiguous() and t.dim() == 5:
        t = t.permute(0, 2, 3, 4, 1)
    return t


@torch.no_grad()
def reduce_tv_step(opt, tv=None):
    """Tail of a ray-sharded training step: gradient all-reduce (mean over ranks) -> total variation -> MaskedAdam, pipelined
    per slab.  The result equals the single-process step on the concatenated batch when the caller's TV weights use the GLOBAL
    ray count (weight / (world * rays_per_rank), as run_train.py:283-287 would with the whole batch on one GPU).

    ``tv`` maps a grid parameter to ``(wx, wy, wz, dense_mode)`` (the arguments of ``total_variation_add_grad``).
    Single process: exactly ``total_variation_add_grad`` on every listed grid followed by ``opt.step()``.
    Several ranks: every slab ``grid[p]`` of a [P,C,X,Y,Z] grid is one all-reduce issued up front on the collective
    stream; the compute stream then waits slab by slab and runs TV + Adam on slab p while slabs p+1.. are still on the
    wire.  This is legal because neither sweep couples slabs: TV differences stay inside a slab
    (total_variation_kernel.cu:22-33 -- the leading dims are peeled off by the modulo chain) and Adam is elementwise,
    so the result equals all-reduce-everything -> TV -> step() bit for bit.  Single-slab grids (DenseGrid) are reduced
    whole; the smaller density grid goes first so its sweeps hide behind the k0 transfer."""
    from . import ops
    tv = tv or {}
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    work = []                                             # (group, param, slice-or-None)
    for group in opt.param_groups:
        group['skip_zero_grad']
        for param in group['params']:
            if param.grad is None:
                continue
            _normalise_grad_layout(param)
            if world > 1 and param.dim() == 5 and param.shape[0] > 1:
                work.extend((group, param, slice(p, p + 1)) for p in range(param.shape[0]))
            else:
                work.append((group, param, None))
    work.sort(key=lambda w: w[1].numel())                 # stable: small tensors first, slabs stay in order
    handles = []
    if world > 1:
        for _, param, sl in work:
            g = param.grad if sl is None else param.grad[sl]
            handles.append(_MeanReduce(_memory_order(g)))
    states = {}
    for i, (group, param, sl) in enumerate(work):
        if handles:
            handles[i].wait()
        if param in tv:
            wx, wy, wz, dense = tv[param]
            if sl is None:
                ops.total_variation_add_grad(param, param.grad, wx, wy, wz, dense)
            else:
                ops.total_variation_add_grad(param[sl], param.grad[sl], wx, wy, wz, dense)
        if id(param) not in states:
            states[id(param)] = opt._begin(param)
        opt._apply(group, param, states[id(param)], sl)


# ----------------------------------------------------------------------------------------------------------------------
# Multi-GPU tail over NVLink peer memory: reduce-scatter -> TV -> Adam -> all-gather in one sweep (csrc/grid_sweep.cu,
# k_tv_adam_peer).  Replaces "all-reduce 1.7 GB of gradient, then every rank repeats the full TV + Adam sweeps".
# ----------------------------------------------------------------------------------------------------------------------
def peer_mapped_buffers(numel, device, count):
    """``count`` fp32 buffers of ``numel`` elements on every rank, each mapped into every other rank's address space.
    Returns [(local flat tensor, [device address of that buffer on rank 0..world-1, valid in THIS process], keepalive), ...].
    Collective (same call order on all ranks).  Mapping: torch symmetric memory (CUDA VMM handles exchanged by the c10d store);
    fallback: CUDA IPC handles of ordinary allocations (the torch.multiprocessing tensor-sharing mechanism).  world == 1: plain
    allocations."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    out = []
    if world == 1:
        for _ in range(count):
            t = torch.zeros(numel, dtype=torch.float32, device=device)
            out.append((t, [t.data_ptr()], None))
        return out
    mode = os.environ.get('UBN_PEER_MAP', 'auto')
    if mode in ('auto', 'symm'):
        try:
            import torch.distributed._symmetric_memory as symm_mem
            made = []
            for _ in range(count):
                t = symm_mem.empty(numel, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(t, dist.group.WORLD)
                t.zero_()
                made.append((t, [int(a) for a in hdl.buffer_ptrs], hdl))
            return made
        except Exception as e:                       # pragma: no cover  (depends on the box)
            if mode == 'symm':
                raise
            print(f'[ubn.dist] symmetric memory unavailable ({e!r}); falling back to CUDA IPC', flush=True)
    rank = dist.get_rank()
    for _ in range(count):
        t = torch.zeros(numel, dtype=torch.float32, device=device)
        meta = (t.untyped_storage()._share_cuda_(), t.storage_offset() * t.element_size())
        metas = [None] * world
        dist.all_gather_object(metas, meta)
        ptrs, keep = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(t.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*metas[r][0])
            keep.append(st)
            ptrs.append(st.data_ptr() + metas[r][1])
        out.append((t, ptrs, keep))
    return out


class _PeerGrid:
    """A replicated channels-last grid parameter with its gradient buffer and both ping-pong parameter buffers peer-mapped."""

    def __init__(self, param, rank, world):
        P, C, X, Y, Z = param.shape
        n = param.numel()
        (self.g, self.g_ptrs, k0), (self.a, self.a_ptrs, k1), (self.b, self.b_ptrs, k2) = peer_mapped_buffers(n, param.device, 3)
        self._keep = (k0, k1, k2)
        view = lambda flat: flat.view(P, X, Y, Z, C).permute(0, 4, 1, 2, 3)
        self.view = view
        with torch.no_grad():
            view(self.a).copy_(param.detach())
            param.data = view(self.a)
        param._ubn_grad_buffer = view(self.g)
        self.param = param
        self.planes = P * X
        self.lo, self.hi = shard_range(self.planes, rank, world)

    def swap(self):
        self.a, self.b = self.b, self.a
        self.a_ptrs, self.b_ptrs = self.b_ptrs, self.a_ptrs
        self.param.data = self.view(self.a)


class PeerTail:
    """Tail of a ray-sharded training step with the big grids exchanged over NVLink peer memory.

        tail = PeerTail(opt)                 # once, after the optimizer exists (collective); moves eligible grids to peer memory
        loss.backward()                      # the march scatters straight into the persistent peer-visible gradient buffers
        tail.step(tv_terms)                  # == mean-all-reduce -> total_variation_add_grad -> opt.step(), see below

    Per eligible grid (channels-last [P,C,X,Y,Z], C % 4 == 0, no per-voxel lr) rank r owns planes shard_range(P*X, r, world) of
    the flattened (slab, X) axis and runs ONE kernel on them: mean of all ranks' gradients (P2P loads) -> TV -> (masked) Adam on
    its own moments -> the updated parameters stored into every rank's spare parameter buffer (P2P stores).  Parameters
    ping-pong, so nobody overwrites values a neighbour's TV stencil still reads.  Two cross-rank barriers bracket the sweeps
    (gradients complete / stores complete); then each rank zeroes its gradient buffer and swaps the parameter buffers.  The
    sweep work and the Adam state traffic divide by world, each NVLink direction carries (world-1)/world of the grid once.
    Everything else (C == 1 density grids, the rgbnet) takes the classic route: mean all-reduce over NCCL + TV + MaskedAdam.
    Exactness: identical to reduce_tv_step up to the fp32 summation order of the ranks' gradients (fixed: rank order);
    world == 1 is bit-identical to total_variation_add_grad + opt.step().  Adam moments of a peer grid are only maintained for
    the owned planes (use gather_moments() before saving an optimizer state dict)."""

    def __init__(self, opt):
        from . import ops
        self.opt = opt
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world not in (1, 2, 4, 8):
            raise RuntimeError('PeerTail supports 1, 2, 4 or 8 ranks')
        self.grids = {}
        for group in opt.param_groups:
            for param in group['params']:
                if opt.per_lr is None and param.requires_grad and ops.tv_adam_pingpong_supported(param):
                    self.grids[param] = _PeerGrid(param, self.rank, self.world)
        self._flag = None
        if self.world > 1:
            self._flag = torch.zeros(1, device=next(iter(self.grids)).device if self.grids else 'cuda')
            dist.barrier()

    def _barrier(self):
        """Stream-ordered cross-rank barrier: a 4-byte NCCL all-reduce.  It starts on a rank once that rank's stream has reached
        this point and completes nowhere before every rank has joined."""
        if self.world > 1:
            dist.all_reduce(self._flag)

    @torch.no_grad()
    def step(self, tv=None):
        from . import ops
        tv = tv or {}
        opt = self.opt
        classic = []
        for group in opt.param_groups:
            group['skip_zero_grad']
            for param in group['params']:
                if param in self.grids:
                    if param.grad is None:                                  # not touched this step: still consume a zero gradient
                        param.grad = param._ubn_grad_buffer
                    continue
                if param.grad is not None:
                    _normalise_grad_layout(param)
                    classic.append((group, param))
        classic.sort(key=lambda w: w[1].numel())
        handles = [_MeanReduce(_memory_order(param.grad)) for _, param in classic] if self.world > 1 else []
        self._barrier()                                                     # every rank's gradient buffers are complete
        for group in opt.param_groups:
            for param in group['params']:
                pg = self.grids.get(param)
                if pg is None:
                    continue
                state = opt._begin(param)
                wx, wy, wz, dense = tv.get(param, (0.0, 0.0, 0.0, True))
                beta1, beta2 = group['betas']
                ops.tv_adam_peer(param, pg.b_ptrs, pg.g_ptrs, state['exp_avg'], state['exp_avg_sq'], wx, wy, wz, dense, pg.lo, pg.hi,
                                 state['step'], beta1, beta2, group['lr'], group['eps'], skip_zero_grad=group['skip_zero_grad'])
        self._barrier()                                                     # every rank's parameter stores have landed
        for pg in self.grids.values():
            pg.g.zero_()
            pg.swap()
            pg.param.grad = None
        for i, (group, param) in enumerate(classic):
            if handles:
                handles[i].wait()
            if param in tv:
                ops.total_variation_add_grad(param, param.grad, *tv[param])
            opt._apply(group, param, opt._begin(param))

    @torch.no_grad()
    def gather_moments(self):
        """Make exp_avg / exp_avg_sq of the peer grids whole on every rank (each rank maintains only its owned planes): one
        all-reduce of the moments masked to the owned range.  For checkpointing; not part of the step."""
        if self.world == 1:
            return
        for param, pg in self.grids.items():
            st = self.opt.state.get(param, {})
            for k in ('exp_avg', 'exp_avg_sq'):
                if k in st:
                    flat = _memory_order(st[k]).reshape(pg.planes, -1)
                    flat[:pg.lo].zero_()
                    flat[pg.hi:].zero_()
                    dist.all_reduce(flat)


def gather_frame(local_out, n_total, rank, world, dst=0):
    """Render path: every rank rendered its contiguous shard of a frame ([n_local, K] rgb/depth/...) -> all ranks get
    the assembled [n_total, K] tensor (one all_gather of at most ceil(n_total/world) rows per rank)."""
    if world == 1:
        return local_out
    per = (n_total + world - 1) // world
    pad = torch.zeros(per, *local_out.shape[1:], dtype=local_out.dtype, device=local_out.device)
    pad[:local_out.shape[0]] = local_out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(bufs[r][:hi - lo])
    return torch.cat(parts, 0)


def idw_composite(rgb, origins, centroid, power=4):
    """Block-NeRF style compositing of per-block renders (eval_block_nerf.py:95-98,123-127): every rank holds one block,
    renders the same rays, and the frame is sum_b w_b rgb_b / sum_b w_b with w_b = ||o - c_b||^-p (all-reduce SUM)."""
    w = (origins - centroid).norm(dim=-1, keepdim=True).clamp_min(1e-8).pow(-power)
    num = rgb * w
    den = w.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(num)
        dist.all_reduce(den)
    return num / den
