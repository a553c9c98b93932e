"""TEST INFRASTRUCTURE (oracle side): "PSNR delta vs ref", the second half of BASELINE.json's metric, without datasets.

Protocol of SURVEY.md 8(d): a seeded smooth procedural scene (the teacher) rendered by the ORACLE to two pinhole views is
the ground truth; the student = teacher + seeded noise is rendered once by the oracle (the reference's arithmetic,
oracle/cpu_ref.py) and once by the CUDA library on identical rays and weights;
    delta = PSNR(cuda, GT) - PSNR(oracle, GT)     (gate: |delta| <= 0.01 dB)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this (as the checker)."""
import torch

from . import cpu_ref


def pinhole_rays(H, W, cam_pos, look_at=(0., 0., 0.), focal=None):
    """A small pinhole view -> (rays_o, rays_d, viewdirs), pixel centres, z-up."""
    focal = focal or float(W)
    cam_pos = torch.tensor(cam_pos)
    fwd = torch.tensor(look_at) - cam_pos
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0., 0., 1.]))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing='ij')
    d = ((i - W / 2) / focal)[..., None] * right + (-(j - H / 2) / focal)[..., None] * up + fwd
    rd = d.reshape(-1, 3).contiguous()
    ro = cam_pos.expand_as(rd).contiguous()
    return ro, rd, (rd / rd.norm(dim=-1, keepdim=True)).contiguous()


def psnr(a, b):
    return -10.0 * torch.log10(((a.double() - b.double()) ** 2).mean()).item()


def model_kwargs(flavor, world, F_, thres=1e-4):
    if flavor == 'fouriergrid':
        return dict(xyz_min=[-1.] * 3, xyz_max=[1.] * 3, num_voxels_density=world ** 3, num_voxels_base_density=world ** 3,
                    num_voxels_rgb=world ** 3, num_voxels_base_rgb=world ** 3, num_voxels_viewdir=-1, alpha_init=1e-4,
                    fast_color_thres=thres, rgbnet_dim=12, fourier_freq_num=F_, contracted_norm='inf')
    return dict(xyz_min=[-1.] * 3, xyz_max=[1.] * 3, num_voxels=world ** 3, num_voxels_base=world ** 3, alpha_init=1e-4,
                fast_color_thres=thres, rgbnet_dim=12, contracted_norm='inf')


@torch.no_grad()
def paint_teacher(m, flavor, g):
    """Smooth procedural scene: low-pass noise + a dense ball of radius 0.55 (FourierGrid averages 1+2F warped look-ups, so
    the scene reads as lumpy fog around act_shift = -9.2 rather than a solid)."""
    def smooth(shape, amp):
        low = torch.randn(shape[0], shape[1], 6, 6, 6, generator=g)
        return torch.nn.functional.interpolate(low, size=shape[2:], mode='trilinear', align_corners=True) * amp
    X = m.density.grid.shape[2]
    ax = torch.linspace(-1.2, 1.2, X)
    rr = (ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2).sqrt()
    ball = ((0.55 - rr) * 40.0).clamp(-4.0, 8.0)
    m.density.grid.copy_(smooth(m.density.grid.shape, 3.0) + ball[None, None] + 7.0)
    m.k0.grid.copy_(smooth(m.k0.grid.shape, 1.5))
    if flavor == 'dcvgo':
        m.mask_cache.mask.fill_(True)


def psnr_delta(flavor, F_, device, world=32, H=24, W=24, seed=4242):
    """Returns dict(psnr_oracle, psnr_cuda, delta_db, psnr_cuda_vs_oracle, gt_std).  `device` runs the library under test."""
    from unboundednerfpytorch_b200 import models
    torch.manual_seed(seed)
    kw = model_kwargs(flavor, world, F_)
    m = (models.FourierGridModel if flavor == 'fouriergrid' else models.DirectContractedVoxGO)(**kw)
    g = torch.Generator().manual_seed(99)
    paint_teacher(m, flavor, g)
    views = [pinhole_rays(H, W, (2.2, 0.3, 0.4)), pinhole_rays(H, W, (-0.5, -2.0, 1.0))]

    def render_oracle(state):
        p = cpu_ref.params_from_state(flavor, kw, state)
        with torch.no_grad():
            return torch.cat([cpu_ref.model_forward(flavor, p, *v, 0.5, bg=1, render_depth=False)['rgb_marched'] for v in views])

    snap = lambda: {k: v.detach().clone().contiguous() for k, v in m.state_dict().items()}
    gt = render_oracle(snap())
    with torch.no_grad():                                                 # student = teacher + noise
        m.density.grid.add_(torch.randn(m.density.grid.shape, generator=g) * 0.5)
        m.k0.grid.add_(torch.randn(m.k0.grid.shape, generator=g) * 0.3)
    img_ref = render_oracle(snap())
    m = m.to(device)
    with torch.no_grad():
        img_new = torch.cat([m(*(t.to(device) for t in v), global_step=None, is_train=False, near=0., far=1e9, bg=1,
                               rand_bkgd=False, stepsize=0.5, render_depth=False)['rgb_marched'].cpu() for v in views])
    a, b = psnr(img_ref, gt), psnr(img_new, gt)
    return dict(psnr_oracle=a, psnr_cuda=b, delta_db=b - a, psnr_cuda_vs_oracle=psnr(img_new, img_ref), gt_std=gt.std().item())
