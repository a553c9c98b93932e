"""CPU: pin the oracle (oracle/ref_ops.c + oracle/cpu_ref.py) against fixtures produced by the REFERENCE'S OWN
Python files run in the build container (oracle/make_golden.py).  The reference itself ships no tests or golden
vectors, so these fixtures are the pin."""
import torch

from tests.util import assert_close, assert_equal, load_golden


def test_dense_grid_oracle_matches_reference(oracle):
    g = load_golden('l1_grids.pt')
    for key in ('dense_C1', 'dense_C3', 'dense_C12'):
        r = g[key]
        grid = r['grid'].clone().requires_grad_(True)
        out = oracle.dense_grid_forward(grid, r['xyz'], torch.tensor(r['xyz_min']), torch.tensor(r['xyz_max']))
        assert_close(out, r['out'], what=key)
        (out * r['w']).sum().backward()
        assert_close(grid.grad, r['grad_grid'], what=key + ' grad')


def test_fourier_grid_oracle_matches_reference(oracle):
    g = load_golden('l1_grids.pt')
    for key in ('fourier_C1_F2', 'fourier_C12_F4', 'fourier_C3_F1'):
        r = g[key]
        grid = r['grid'].clone().requires_grad_(True)
        out = oracle.fourier_grid_forward(grid, r['xyz'], torch.tensor(r['xyz_min']), torch.tensor(r['xyz_max']), r['num_freqs'])
        assert_close(out, r['out'], what=key)
        (out * r['w']).sum().backward()
        assert_close(grid.grad, r['grad_grid'], what=key + ' grad')


def test_maskgrid_tv_oracle(oracle):
    g = load_golden('l1_grids.pt')
    r = g['maskgrid']
    assert_equal(oracle.maskcache_lookup(r['mask'], r['xyz'], r['scale'], r['shift']), r['out'], 'maskcache')
    tv = g['tv']
    for k in ('dense1', 'dense0'):
        grad = tv[k]['grad_in'].clone()
        oracle.total_variation_add_grad(tv['param'], grad, tv['wx'], tv['wy'], tv['wz'], k == 'dense1')
        assert_close(grad, tv[k]['grad_out'], what='tv ' + k)
    # independent restatement of the TV term with torch slicing (clamped L1 gradient, i-axis uses wz: sic)
    p = tv['param']
    wy, wz = tv['wy'] / 6, tv['wz'] / 6
    add = torch.zeros_like(p)
    for dim, w in ((4, wz), (3, wy), (2, wz)):
        d = (p.narrow(dim, 1, p.shape[dim] - 1) - p.narrow(dim, 0, p.shape[dim] - 1)).clamp(-1, 1) * w
        add.narrow(dim, 1, p.shape[dim] - 1).add_(d)
        add.narrow(dim, 0, p.shape[dim] - 1).sub_(d)
    assert_close(tv['dense1']['grad_out'], tv['dense1']['grad_in'] + add, rtol=1e-5, atol=1e-6, what='tv torch restatement')


def test_autograd_fn_chain_oracle(oracle):
    g = load_golden('l1_autograd_fns.pt')
    r = g['chain']
    exp_d, alpha = oracle.raw2alpha(r['density'], r['shift'], r['interval'])
    assert_close(alpha, r['alpha'], what='alpha')
    w, T, last, i_s, i_e = oracle.alpha2weight(alpha, r['ray_id'], r['n_rays'])
    assert_close(w, r['weights'], what='weights')
    assert_close(last, r['alphainv_last'], what='alphainv_last')
    # early stop happened on the opaque ray and nowhere else
    lens = torch.bincount(r['ray_id'], minlength=r['n_rays'])
    assert (i_e - i_s)[5] < lens[5] and torch.equal((i_e - i_s)[:5], lens[:5])
    ga = oracle.alpha2weight_backward(alpha, w, T, last, i_s, i_e, r['n_rays'], r['gw'], r['gl'])
    gd = oracle.raw2alpha_backward(exp_d, ga, r['interval'])
    assert_close(gd, r['grad_density'], what='grad_density')
    # closed-form check of the forward recurrence in float64 on a non-stopping ray
    seg = r['ray_id'] == 2
    a = alpha[seg].double()
    Tref = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.double), 1 - a[:-1]]), 0)
    assert_close(w[seg], (Tref * a).float(), rtol=1e-5, what='closed form')
    n = g['nonuni']
    e2, a2 = oracle.raw2alpha_nonuni(n['density'], n['shift'], n['interval'])
    assert_close(a2, n['alpha'], what='nonuni alpha')
    assert_close(oracle.raw2alpha_nonuni_backward(e2, n['g'], n['interval']), n['grad_density'], what='nonuni grad')


def test_masked_adam_oracle(oracle):
    g = load_golden('l1_masked_adam.pt')
    for mode, code in (('plain', 0), ('masked', 1), ('perlr', 2)):
        r = g[mode]
        p, q = r['p0'].clone(), r['q0'].clone()
        st = [torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(q), torch.zeros_like(q)]
        for step, ((gp, gq), p_ref, q_ref) in enumerate(zip(r['grads'], r['ps'], r['qs']), 1):
            if code == 2:
                oracle.adam_upd_with_perlr(p, gp, st[0], st[1], r['per_lr'], step, 0.9, 0.99, 0.1, 1e-8)
            elif code == 1:
                oracle.masked_adam_upd(p, gp, st[0], st[1], step, 0.9, 0.99, 0.1, 1e-8)
            else:
                oracle.adam_upd(p, gp, st[0], st[1], step, 0.9, 0.99, 0.1, 1e-8)
            oracle.adam_upd(q, gq, st[2], st[3], step, 0.9, 0.99, 1e-3, 1e-8)
            assert_close(p, p_ref, what=f'{mode} p step {step}')
            assert_close(q, q_ref, what=f'{mode} q step {step}')
            if code == 1:
                assert torch.equal(p[gp == 0], (r['ps'][step - 2] if step > 1 else r['p0'])[gp == 0])
        # torch.optim.Adam restatement (plain mode, eps outside the bias-corrected sqrt like the reference kernel)
        if code == 0:
            b1, b2 = 0.9, 0.99
            p2, m, v = r['p0'].clone().double(), 0, 0
            for step, (gp, _) in enumerate(r['grads'], 1):
                m = b1 * m + (1 - b1) * gp.double()
                v = b2 * v + (1 - b2) * gp.double() ** 2
                ss = 0.1 * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
                p2 = p2 - ss * m / (v.sqrt() + 1e-8)
            assert_close(p, p2.float(), rtol=1e-5, what='adam float64 restatement')


def test_model_forward_oracle_matches_reference(oracle):
    g = load_golden('l2_models.pt')
    for name, flavor in (('fouriergrid_thres', 'fouriergrid'), ('fouriergrid_opaque', 'fouriergrid'),
                         ('dcvgo_inf', 'dcvgo'), ('dcvgo_l2_opaque', 'dcvgo')):
        rec = g[name]
        p = oracle.params_from_state(flavor, rec['kwargs'], rec['state'], requires_grad=True)
        rk, ref = rec['render_kwargs'], rec['ret']
        ret = oracle.model_forward(flavor, p, rec['rays_o'], rec['rays_d'], rec['viewdirs'], rk['stepsize'], bg=rk['bg'])
        assert_equal(ret['ray_id'], ref['ray_id'], name + ' ray_id')
        assert_equal(ret['step_id'], ref['step_id'], name + ' step_id')
        for k in ('rgb_marched', 'alphainv_last', 'weights', 'raw_rgb', 'raw_alpha', 'raw_density', 'depth', 't', 's'):
            assert_close(ret[k], ref[k], what=f'{name} {k}')
        lw = rec['loss_w']
        loss = (ret['rgb_marched'] * lw['rgb']).sum() + (ret['alphainv_last'] * lw['last']).sum()
        loss = loss + 0.01 * (ret['raw_rgb'].pow(2).sum(-1) * ret['weights'].detach()).sum() + 0.1 * ret['weights'].pow(2).sum()
        loss.backward()
        assert_close(p['density_grid'].grad, ref['grads']['density.grid'], what=name + ' d grad')
        assert_close(p['k0_grid'].grad, ref['grads']['k0.grid'], what=name + ' k0 grad')
        assert_close(p['rgbnet']['W2'].grad, ref['grads']['rgbnet.2.0.weight'], what=name + ' W2 grad')


def test_sampling_oracle_matches_reference_dvgo(oracle):
    rec = load_golden('l2_models.pt')['dvgo']
    xyz_min, xyz_max = rec['state']['xyz_min'], rec['state']['xyz_max']
    out = oracle.sample_pts_on_rays(rec['rays_o'].contiguous(), rec['rays_d'].contiguous(), xyz_min, xyz_max, 0.2, 1e9, rec['stepdist'])
    for a, b, nm in zip(out, rec['sample'], ('pts', 'mask_outbbox', 'ray_id', 'step_id', 'N_steps', 't_min', 't_max')):
        (assert_close if a.dtype == torch.float32 else assert_equal)(a, b, what=nm)
    # each ray gets ceil(segment * |d| / stepdist) >= 1 samples and the sample positions advance by stepdist
    n_steps = out[4]
    assert (n_steps >= 1).all() and int(n_steps.sum()) == out[0].shape[0]


def test_rays_oracle_matches_reference():
    """oracle.get_rays_of_a_view == the reference's dvgo.get_rays_of_a_view (dvgo.py:492-557) on every flag combination."""
    from oracle import cpu_ref
    rec = load_golden('l1_rays.pt')
    assert len(rec['views']) == 32
    for v in rec['views']:
        o, d, vd = cpu_ref.get_rays_of_a_view(v['H'], v['W'], v['K'], v['c2w'], v['ndc'], v['inverse_y'], v['flip_x'],
                                              v['flip_y'], mode=v['mode'])
        tag = f"ndc={v['ndc']} inv={v['inverse_y']} fx={v['flip_x']} fy={v['flip_y']} {v['mode']}"
        assert_close(o, v['rays_o'], rtol=1e-6, what='rays_o ' + tag)
        assert_close(d, v['rays_d'], rtol=1e-6, what='rays_d ' + tag)
        assert_close(vd, v['viewdirs'], rtol=1e-6, what='viewdirs ' + tag)
