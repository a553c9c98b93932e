"""CPU: host-side logic of the package (no kernels run): schedules, layouts, descriptors, module surface,
state-dict compatibility with the reference, error behaviour, and the no-fallback rule."""
import os
import re

import pytest
import torch

from tests.util import ROOT, assert_close, load_golden


def test_t_schedule_matches_oracle_and_sizes(oracle):
    from unboundednerfpytorch_b200 import march
    for world_len, S in ((153, 512), (200, 668), (320, 1068)):
        for tb in (1.5, 2.0):
            t = march.t_schedule(world_len, 0.5, 0.2, tb, 'cpu')
            assert t.numel() == S
            assert torch.equal(t, oracle.contracted_t_schedule(world_len, 0.5, 0.2, tb))
            assert (t[1:] > t[:-1]).all()


def test_make_cfg_fields():
    from unboundednerfpytorch_b200 import march
    import numpy as np
    c = march.make_cfg([0, 0, 0], [1, 1, 1], 0.2, 'l2', 512, -9.21, 0.5, 1e-4, cumdist_thres=0.01,
                       mask=torch.ones(4, 5, 6, dtype=torch.bool), mask_scale=[1, 2, 3], mask_shift=[4, 5, 6])
    assert c.contracted_norm == 1 and c.n_samples == 512 and c.use_cumdist == 1 and c.use_maskcache == 1
    assert list(c.mask_sz) == [4, 5, 6]
    assert c.contract_B == float(np.float32(1.2)) and c.contract_A == float(np.float32(0.2))
    with pytest.raises(NotImplementedError):
        march.make_cfg([0] * 3, [1] * 3, 0.2, 'l1', 8, 0, 0.5, 0)


def test_grid_layout_and_desc():
    from unboundednerfpytorch_b200 import grid as G
    g = G.zeros_grid([9, 12, 5, 6, 7])
    assert tuple(g.shape) == (9, 12, 5, 6, 7)
    assert g.stride() == (5 * 6 * 7 * 12, 1, 6 * 7 * 12, 7 * 12, 12)          # channels-last
    assert g.is_contiguous(memory_format=torch.channels_last_3d)
    d = G.grid_desc(g, [-1.2] * 3, [1.2] * 3, 4)
    assert (d.P, d.C, d.X, d.Y, d.Z, d.num_freqs) == (9, 12, 5, 6, 7, 4)
    assert (d.stride_p, d.stride_c, d.stride_v) == (5 * 6 * 7 * 12, 1, 12)
    ref = torch.zeros(1, 3, 5, 6, 7)                                             # reference layout also accepted
    d = G.grid_desc(ref, [-1] * 3, [1] * 3, 0)
    assert (d.stride_p, d.stride_c, d.stride_v) == (3 * 210, 210, 1)
    with pytest.raises(RuntimeError):
        G.grid_desc(torch.zeros(1, 3, 5, 7, 6).transpose(3, 4), [-1] * 3, [1] * 3, 0)   # Y/Z swapped in memory
    # channels-last conversion keeps values
    x = torch.randn(2, 4, 3, 3, 3)
    y = G._as_cl3d(x)
    assert torch.equal(x, y) and y.stride()[1] == 1


def test_models_accept_reference_state_dicts():
    """state_dict keys / shapes / get_kwargs keys of the reference load into the mirror classes unchanged."""
    from unboundednerfpytorch_b200 import models
    g = load_golden('l2_models.pt')
    for name, cls in (('fouriergrid_thres', models.FourierGridModel), ('dcvgo_inf', models.DirectContractedVoxGO)):
        rec = g[name]
        m = cls(**rec['kwargs'])
        missing, unexpected = m.load_state_dict(rec['state'], strict=True)
        assert not missing and not unexpected
        assert_close(m.density.grid, rec['state']['density.grid'], what='density.grid')
        assert m.k0.grid.stride()[1] == 1                                      # still channels-last after loading
        sd = m.state_dict()
        assert sorted(sd.keys()) == sorted(rec["state"].keys())
        for k in sd:
            assert sd[k].shape == rec['state'][k].shape, k
        # round trip through the reference's checkpoint format (FourierGrid_ckpt_manager.py:44-51)
        kw = m.get_kwargs()
        for k in rec['kwargs']:
            if k not in ('fast_color_thres',):
                assert k in kw, k
        m2 = cls(**kw)
        m2.load_state_dict(sd)


def test_masked_adam_surface():
    from unboundednerfpytorch_b200.masked_adam import MaskedAdam, create_optimizer_or_freeze_model
    from unboundednerfpytorch_b200 import models
    lin = torch.nn.Linear(3, 4)
    opt = MaskedAdam(lin.parameters())
    lin.weight.grad = torch.zeros_like(lin.weight)
    with pytest.raises(KeyError):                     # reference quirk: group lacks 'skip_zero_grad' (masked_adam.py:49)
        opt.step()
    with pytest.raises(ValueError):
        MaskedAdam(lin.parameters(), lr=-1)
    rec = load_golden('l2_models.pt')['dcvgo_inf']
    m = models.DirectContractedVoxGO(**rec['kwargs'])
    cfg = dict(lrate_density=1e-1, lrate_k0=1e-1, lrate_rgbnet=1e-3, lrate_decay=20, skip_zero_grad_fields=['density', 'k0'],
               lrate_nonexistent=1.0)
    opt = create_optimizer_or_freeze_model(m, cfg, global_step=0)
    flags = {len(list(g['params'])): g['skip_zero_grad'] for g in opt.param_groups}
    assert [g['skip_zero_grad'] for g in opt.param_groups] == [True, True, False]
    assert opt.param_groups[2]['lr'] == 1e-3
    cfg['lrate_rgbnet'] = 0
    create_optimizer_or_freeze_model(m, cfg, global_step=0)
    assert all(not p.requires_grad for p in m.rgbnet.parameters())


def test_legacy_module_surface():
    import unboundednerfpytorch_b200 as U
    mods = U.install_legacy_modules()
    import render_utils_cuda, total_variation_cuda, adam_upd_cuda, ub360_utils_cuda   # noqa: E401
    want = {'render_utils_cuda': ['infer_t_minmax', 'infer_n_samples', 'infer_ray_start_dir', 'sample_pts_on_rays',
                                  'sample_ndc_pts_on_rays', 'sample_bg_pts_on_rays', 'maskcache_lookup', 'raw2alpha',
                                  'raw2alpha_backward', 'raw2alpha_nonuni', 'raw2alpha_nonuni_backward', 'alpha2weight',
                                  'alpha2weight_backward'],
            'total_variation_cuda': ['total_variation_add_grad'],
            'adam_upd_cuda': ['adam_upd', 'masked_adam_upd', 'adam_upd_with_perlr'],
            'ub360_utils_cuda': ['cumdist_thres']}
    for mod, fns in want.items():
        for f in fns:
            assert callable(getattr(mods[mod], f)), (mod, f)


def test_cpu_tensors_are_rejected_like_the_reference():
    """CHECK_CUDA semantics (render_utils.cpp:46-48): RuntimeError '<name> must be a CUDA tensor'; never a CPU fallback."""
    from unboundednerfpytorch_b200 import ops, grid as G
    with pytest.raises(RuntimeError, match='density must be a CUDA tensor'):
        ops.raw2alpha(torch.zeros(4), 0.0, 0.5)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        ops.alpha2weight(torch.zeros(4), torch.zeros(4, dtype=torch.long), 2)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        ops.cumdist_thres(torch.zeros(2, 3), 0.1)
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        G.DenseGrid(1, [4, 4, 4], [-1] * 3, [1] * 3)(torch.zeros(5, 3))
    with pytest.raises(RuntimeError, match='must be a CUDA tensor'):
        p = torch.zeros(1, 1, 2, 2, 2)
        ops.adam_upd(p, p, p, p, 1, 0.9, 0.99, 0.1, 1e-8)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, 'unboundednerfpytorch_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
                assert 'libubn_oracle' not in src and 'cpu_ref' not in src, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from unboundednerfpytorch_b200 import _cabi
    monkeypatch.setattr(_cabi, '_lib', None)
    monkeypatch.setattr(_cabi, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU or eager-PyTorch fallback'):
        _cabi.load()


def test_rays_host_side_pieces():
    """rays.py parts that are plain torch / numpy (the kernels are covered by the -m gpu tests): ndc_rays == the oracle's
    restatement of dvgo.py:532-550 on the golden view, batch_indices_generator covers every index once per epoch."""
    import numpy as np
    from oracle import cpu_ref
    from tests.util import assert_close, load_golden
    from unboundednerfpytorch_b200 import rays as R
    v = [x for x in load_golden('l1_rays.pt')['views'] if x['ndc'] and not x['flip_x'] and not x['flip_y'] and x['mode'] == 'center'][0]
    o, d, _ = cpu_ref.get_rays_of_a_view(v['H'], v['W'], v['K'], v['c2w'], False, v['inverse_y'], False, False)
    o2, d2 = R.ndc_rays(v['H'], v['W'], float(v['K'][0][0]), 1., o, d)
    assert_close(o2, v['rays_o'], rtol=1e-6, what='ndc rays_o'); assert_close(d2, v['rays_d'], rtol=1e-6, what='ndc rays_d')
    np.random.seed(0)
    gen = R.batch_indices_generator(10, 4)
    first_epoch = torch.cat([next(gen) for _ in range(2)])
    assert first_epoch.dtype == torch.int64 and len(set(first_epoch.tolist())) == 8
    nxt = next(gen)                                   # 8 + 4 > 10 -> reshuffle, like dvgo.py:663-665
    assert nxt.shape == (4,) and int(nxt.max()) < 10
    with pytest.raises(NotImplementedError):
        R._rays_of_a_view(2, 2, np.eye(3), np.eye(4)[:3], False, False, False, False, 'bogus')


def test_reference_checkpoint_interchange(tmp_path):
    """a checkpoint written by the REFERENCE's classes (tests/golden/ref_fine_last.tar, produced by
    oracle/make_golden.py::golden_checkpoint with FourierGrid_model.py + masked_adam.py) loads into this library's model and
    optimizer; a checkpoint written here has the reference's keys, contiguous [P,C,X,Y,Z] grids, and loads back unchanged."""
    import os
    import numpy as np
    import torch
    from unboundednerfpytorch_b200 import ckpt, models
    from unboundednerfpytorch_b200.masked_adam import MaskedAdam
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_fine_last.tar')
    raw = torch.load(path, map_location='cpu', weights_only=False)
    assert set(raw) == {'global_step', 'model_kwargs', 'model_state_dict', 'optimizer_state_dict'}
    assert isinstance(raw['model_kwargs']['xyz_min'], np.ndarray)            # why weights_only=True cannot load reference files
    m = ckpt.load_model(models.FourierGridModel, path)
    for k, v in raw['model_state_dict'].items():
        assert torch.equal(m.state_dict()[k].contiguous(), v), k
    assert m.k0.grid.stride()[1] == 1                                        # copied into the channels-last layout
    opt = MaskedAdam([{'params': [m.density.grid], 'lr': 0.1, 'skip_zero_grad': True},
                      {'params': [m.k0.grid], 'lr': 0.1, 'skip_zero_grad': True},
                      {'params': list(m.rgbnet.parameters()), 'lr': 1e-3, 'skip_zero_grad': False}])
    m2, opt2, start = ckpt.load_checkpoint(m, opt, path, no_reload_optimizer=False)
    assert start == 2 and opt2.state[m.k0.grid]['step'] == 2
    assert opt2.state[m.k0.grid]['exp_avg'].stride() == m.k0.grid.stride()
    assert torch.equal(opt2.state[m.k0.grid]['exp_avg'].contiguous(), raw['optimizer_state_dict']['state'][1]['exp_avg'])
    out = tmp_path / 'fine_last.tar'
    ckpt.save_checkpoint(7, m, opt, str(out))
    back = torch.load(str(out), map_location='cpu', weights_only=False)
    assert back['global_step'] == 7 and set(back) == set(raw)
    assert back['model_state_dict']['k0.grid'].is_contiguous()
    for k, v in raw['model_state_dict'].items():
        assert torch.equal(back['model_state_dict'][k], v), k
    assert sorted(back['model_kwargs']) == sorted(raw['model_kwargs'])
