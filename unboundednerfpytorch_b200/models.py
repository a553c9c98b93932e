"""Unbounded-scene voxel-grid radiance-field models with the reference's constructor / forward surface:

* ``FourierGridModel``       -- FourierGrid/FourierGrid_model.py:134-681
* ``DirectContractedVoxGO``  -- FourierGrid/dcvgo.py:28-384

Same constructor keywords, ``get_kwargs()`` keys, state-dict names (``density.grid``, ``k0.grid``,
``rgbnet.{0,2.0,3}.{weight,bias}``, ``mask_cache.mask`` ...), ``forward(rays_o, rays_d, viewdirs,
global_step=None, is_train=False, **render_kwargs)`` and ``ret_dict`` keys, so run_train.py /
run_render.py style callers work unchanged.  ``forward`` runs the fused march (march.py: 3 launches
instead of ~40 and no boolean-mask compaction syncs except the one ragged-size read); ``forward_ops``
composes the individual drop-in ops in the reference's order (used for cross-checks and for grids the
fused feature kernel does not cover, e.g. the 3-channel coarse-stage k0).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import grid as G
from . import march
from . import shade as shade_mod
from .functional import Alphas2Weights, Raw2Alpha, composite_rgb, host_scalar, segment_sum


def _cube_root_size(xyz_min, xyz_max, num_voxels):
    return ((xyz_max - xyz_min).prod() / num_voxels).pow(1 / 3)


def _make_rgbnet(dim0, width, depth):
    # Linear-ReLU-(Sequential(Linear,ReLU))*-Linear: identical module tree => identical state-dict keys
    net = nn.Sequential(
        nn.Linear(dim0, width), nn.ReLU(inplace=True),
        *[nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True)) for _ in range(depth - 2)],
        nn.Linear(width, 3))
    nn.init.constant_(net[-1].bias, 0)
    return net


def _view_embed(viewdirs, viewfreq):
    emb = (viewdirs.unsqueeze(-1) * viewfreq).flatten(-2)
    return torch.cat([viewdirs, emb.sin(), emb.cos()], -1)


class _ContractedBase(nn.Module):
    """Shared machinery of the two contracted-space models."""

    T_BOUNDARY = 2.0          # dcvgo.py:243-244; FourierGridModel overrides with 1.5
    USE_CUMDIST = False
    USE_MASKCACHE = False

    # ---- helpers --------------------------------------------------------------------------------
    def _host(self):
        if self._host_cache is None:
            self._host_cache = dict(center=self.scene_center.detach().cpu().tolist(),
                                    radius=self.scene_radius.detach().cpu().tolist())
        return self._host_cache

    def _apply(self, fn, *a, **k):
        self._host_cache = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._host_cache = None
        return super()._load_from_state_dict(*a, **k)

    def _init_scene(self, xyz_min, xyz_max, bg_len, fast_color_thres, contracted_norm):
        xyz_min = torch.as_tensor(np.asarray(xyz_min), dtype=torch.float32)
        xyz_max = torch.as_tensor(np.asarray(xyz_max), dtype=torch.float32)
        assert len(((xyz_max - xyz_min) * 100000).long().unique()), 'scene bbox must be a cube'
        self.register_buffer('scene_center', (xyz_min + xyz_max) * 0.5)
        self.register_buffer('scene_radius', (xyz_max - xyz_min) * 0.5)
        self.register_buffer('xyz_min', torch.tensor([-1., -1., -1.]) - bg_len)
        self.register_buffer('xyz_max', torch.tensor([1., 1., 1.]) + bg_len)
        if isinstance(fast_color_thres, dict):
            self._fast_color_thres = fast_color_thres
            self.fast_color_thres = fast_color_thres[0]
        else:
            self._fast_color_thres = None
            self.fast_color_thres = fast_color_thres
        self.bg_len = bg_len
        self.contracted_norm = contracted_norm
        self._host_cache = None

    def _maybe_update_thres(self, global_step):
        if isinstance(self._fast_color_thres, dict) and global_step in self._fast_color_thres:
            self.fast_color_thres = self._fast_color_thres[global_step]

    def activate_density(self, density, interval=None):
        interval = interval if interval is not None else self._voxel_size_ratio()
        shape = density.shape
        return Raw2Alpha.apply(density.flatten().contiguous(), self.act_shift, interval).reshape(shape)

    def _sample_dense(self, ori_rays_o, ori_rays_d, stepsize):
        """Dense [N,S,3] contracted points with torch elementwise ops (dcvgo.py:239-262,
        FourierGrid_model.py:522-552) -- only the op-by-op path and API users need them materialised."""
        rays_o = (ori_rays_o - self.scene_center) / self.scene_radius
        rays_d = ori_rays_d / ori_rays_d.norm(dim=-1, keepdim=True)
        t = march.t_schedule(self._world_len(), stepsize, self.bg_len, self.T_BOUNDARY, ori_rays_o.device)
        ray_pts = rays_o[:, None, :] + rays_d[:, None, :] * t[None, :, None]
        if self.contracted_norm == 'inf':
            norm = ray_pts.abs().amax(dim=-1, keepdim=True)
        elif self.contracted_norm == 'l2':
            norm = ray_pts.norm(dim=-1, keepdim=True)
        else:
            raise NotImplementedError
        inner_mask = (norm <= 1)
        B = 1 + self.bg_len
        A = B * 1.0 - 1.0
        ray_pts = torch.where(inner_mask, ray_pts, ray_pts / norm * (B - A / norm))
        return ray_pts, inner_mask.squeeze(-1), t

    # ---- the fused forward --------------------------------------------------------------------------
    def _fused_ok(self):
        kg = self.k0.grid
        return kg.is_cuda and kg.shape[1] in (4, 8, 12, 16) and kg.shape[0] <= 16

    def _march(self, rays_o, rays_d, stepsize, coherent=False):
        dev = rays_o.device
        t_table = march.t_schedule(self._world_len(), stepsize, self.bg_len, self.T_BOUNDARY, dev)
        interval = stepsize * float(self._voxel_size_ratio())
        host = self._host()
        mask = self.mask_cache.mask if self.USE_MASKCACHE else None
        cum = None
        if self.USE_CUMDIST:
            cum = (2 + 2 * self.bg_len) / self._world_len() * stepsize * 0.95
        mscale = mshift = None
        if mask is not None:
            if getattr(self, '_mask_host', None) is None or self._mask_host[0] is not self.mask_cache:
                self._mask_host = (self.mask_cache, self.mask_cache.xyz2ijk_scale.cpu().tolist(),
                                   self.mask_cache.xyz2ijk_shift.cpu().tolist())
            mscale, mshift = self._mask_host[1], self._mask_host[2]
        cfg = march.make_cfg(host['center'], host['radius'], self.bg_len, self.contracted_norm, t_table.numel(),
                             host_scalar(self.act_shift), interval, self.fast_color_thres,
                             cumdist_thres=cum, mask=mask, mask_scale=mscale, mask_shift=mshift)
        dmn, dmx = self.density._bounds()
        kmn, kmx = self.k0._bounds()
        ddesc = G.grid_desc(self.density.grid, dmn, dmx, self.density.num_freqs)
        kdesc = G.grid_desc(self.k0.grid, kmn, kmx, self.k0.num_freqs)
        dense_known = (self.fast_color_thres <= 0) and not self.USE_CUMDIST and not self.USE_MASKCACHE
        out = march.March.apply(self.density.grid, self.k0.grid, rays_o, rays_d, t_table, mask, cfg, ddesc, kdesc,
                                dense_known, coherent)
        return out, t_table

    def _shade(self, k0, viewdirs, ray_id):
        if self.rgbnet is None:
            return torch.sigmoid(k0)
        emb = _view_embed(viewdirs, self.viewfreq).flatten(0, -2)
        if shade_mod.supported(self.rgbnet, k0.shape[-1]):
            return shade_mod.shade(self.rgbnet, k0, emb, ray_id)          # fused on-chip MLP (csrc/shade.cu)
        return torch.sigmoid(self.rgbnet(torch.cat([k0, emb[ray_id]], -1)))   # other widths / depths: torch (cuBLAS)

    def density_total_variation_add_grad(self, weight, dense_mode):
        w = weight * self._tv_world_max(self.density) / 128
        self.density.total_variation_add_grad(w, w, w, dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        w = weight * self._tv_world_max(self.k0) / 128
        self.k0.total_variation_add_grad(w, w, w, dense_mode)

    def tv_terms(self, weight_density=0., weight_k0=0., dense_mode=True):
        """{grid parameter: (wx, wy, wz, dense_mode)} with the weights of the two methods above -- the form
        dist.reduce_tv_step consumes to pipeline all-reduce / TV / Adam slab by slab."""
        out = {}
        for grid, weight in ((self.density, weight_density), (self.k0, weight_k0)):
            if weight > 0:
                w = weight * self._tv_world_max(grid) / 128
                out[grid.grid] = (w, w, w, dense_mode)
        return out

    @staticmethod
    def _tv_world_max(g):
        return float(max(g.grid.shape[2:]))


@torch.no_grad()
def _occupancy_update(model, density, mask_cache, interval):
    """mask &= max_pool3d(Raw2Alpha(density(lattice of the mask grid))) > fast_color_thres  (two launches, 4 B + 1 B per cell)."""
    from . import ops
    from .functional import host_scalar
    mn, mx = density._bounds()
    alpha = ops.lattice_alpha(density.grid.data, mn, mx, density.num_freqs, model.xyz_min.tolist(), model.xyz_max.tolist(),
                              mask_cache.mask.shape, host_scalar(model.act_shift), interval)
    ops.maxpool3_gt_and_(mask_cache.mask, alpha, model.fast_color_thres)


# ======================================================================================================
class FourierGridModel(_ContractedBase):
    """FourierGrid/FourierGrid_model.py:134-681."""

    T_BOUNDARY = 1.5          # FourierGrid_model.py:526

    def __init__(self, xyz_min, xyz_max, num_voxels_density=0, num_voxels_base_density=0, num_voxels_rgb=0,
                 num_voxels_base_rgb=0, num_voxels_viewdir=0, alpha_init=None, mask_cache_world_size=None,
                 fast_color_thres=0, bg_len=0.2, contracted_norm='inf', density_type='DenseGrid', k0_type='DenseGrid',
                 density_config={}, k0_config={}, rgbnet_dim=0, rgbnet_depth=3, rgbnet_width=128, fourier_freq_num=5,
                 viewbase_pe=4, img_emb_dim=-1, verbose=False, **kwargs):
        super().__init__()
        self._init_scene(xyz_min, xyz_max, bg_len, fast_color_thres, contracted_norm)
        self.verbose = verbose
        self.fourier_freq_num = fourier_freq_num
        self.num_voxels_base_density = num_voxels_base_density
        self.voxel_size_base_density = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_base_density)
        self.num_voxels_base_rgb = num_voxels_base_rgb
        self.voxel_size_base_rgb = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_base_rgb)
        self.num_voxels_viewdir = num_voxels_viewdir
        self._set_grid_resolution(num_voxels_density, num_voxels_rgb)
        self.alpha_init = alpha_init
        self.register_buffer('act_shift', torch.FloatTensor([np.log(1 / (1 - alpha_init) - 1)]))
        self.density_type, self.density_config = density_type, density_config
        self.k0_type, self.k0_config = k0_type, k0_config
        self.world_size = self.world_size_density
        self.density = G.FourierGrid(channels=1, world_size=self.world_size_density, xyz_min=self.xyz_min,
                                     xyz_max=self.xyz_max, use_nerf_pos=True, fourier_freq_num=fourier_freq_num,
                                     config=density_config)
        self.rgbnet_kwargs = dict(rgbnet_dim=rgbnet_dim, rgbnet_depth=rgbnet_depth, rgbnet_width=rgbnet_width,
                                  viewbase_pe=viewbase_pe)
        self.sample_num = kwargs.get('sample_num', -1)
        self.img_embeddings, self.img_embed_dim, self.pos_emb = None, 0, None
        if rgbnet_dim <= 0:
            self.k0_dim = 3
            self.k0 = G.FourierGrid(channels=3, world_size=self.world_size_rgb, xyz_min=self.xyz_min,
                                    xyz_max=self.xyz_max, use_nerf_pos=False, fourier_freq_num=fourier_freq_num,
                                    config=k0_config)
            self.rgbnet = None
        else:
            self.k0_dim = rgbnet_dim
            self.k0 = G.FourierGrid(channels=rgbnet_dim, world_size=self.world_size_rgb, xyz_min=self.xyz_min,
                                    xyz_max=self.xyz_max, use_nerf_pos=True, fourier_freq_num=fourier_freq_num,
                                    config=k0_config)
            self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
            self.rgbnet = _make_rgbnet(3 + 3 * viewbase_pe * 2 + rgbnet_dim, rgbnet_width, rgbnet_depth)
        self.vd = None      # view-direction grid variant (num_voxels_viewdir > 0) is not on the benchmarked path
        if num_voxels_viewdir is not None and num_voxels_viewdir > 0:
            raise NotImplementedError('num_voxels_viewdir > 0 (view-direction grid) is outside the hot-path scope')
        if mask_cache_world_size is None:
            mask_cache_world_size = self.world_size_density
        self.mask_cache = G.MaskGrid(path=None, mask=torch.ones([int(v) for v in mask_cache_world_size], dtype=torch.bool),
                                     xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _set_grid_resolution(self, num_voxels_density, num_voxels_rgb):
        self.num_voxels_density, self.num_voxels_rgb = num_voxels_density, num_voxels_rgb
        self.voxel_size_density = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_density)
        self.voxel_size_rgb = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_rgb)
        self.world_size_density = ((self.xyz_max - self.xyz_min) / self.voxel_size_density).long()
        self.world_size_rgb = ((self.xyz_max - self.xyz_min) / self.voxel_size_rgb).long()
        self.world_len_density = self.world_size_density[0].item()
        self.world_len_rgb = self.world_size_rgb[0].item()
        self.voxel_size_ratio_density = self.voxel_size_density / self.voxel_size_base_density
        self.voxel_size_ratio_rgb = self.voxel_size_rgb / self.voxel_size_base_rgb

    def _world_len(self):
        return self.world_len_density

    def _voxel_size_ratio(self):
        return self.voxel_size_ratio_density

    def get_kwargs(self):
        return {
            'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
            'num_voxels_density': self.num_voxels_density, 'num_voxels_rgb': self.num_voxels_rgb,
            'num_voxels_viewdir': self.num_voxels_viewdir, 'fourier_freq_num': self.fourier_freq_num,
            'num_voxels_base_density': self.num_voxels_base_density, 'num_voxels_base_rgb': self.num_voxels_base_rgb,
            'alpha_init': self.alpha_init, 'voxel_size_ratio_density': self.voxel_size_ratio_density,
            'voxel_size_ratio_rgb': self.voxel_size_ratio_rgb,
            'mask_cache_world_size': list(self.mask_cache.mask.shape), 'fast_color_thres': self.fast_color_thres,
            'contracted_norm': self.contracted_norm, 'density_type': self.density_type, 'k0_type': self.k0_type,
            'density_config': self.density_config, 'k0_config': self.k0_config, 'sample_num': self.sample_num,
            **self.rgbnet_kwargs,
        }

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels_density, num_voxels_rgb):
        self._set_grid_resolution(num_voxels_density, num_voxels_rgb)
        self.density.scale_volume_grid(self.world_size_density)
        self.k0.scale_volume_grid(self.world_size_rgb)
        self.world_size = self.world_size_density
        if np.prod(self.world_size_density.tolist()) <= 256 ** 3:
            self._rebuild_mask_cache(self.world_size_density)

    @torch.no_grad()
    def _rebuild_mask_cache(self, world_size):
        dev = self.density.grid.device
        ws = [int(v) for v in world_size]
        axes = [torch.linspace(float(self.xyz_min[a]), float(self.xyz_max[a]), ws[a], device=dev) for a in range(3)]
        xyz = torch.stack(torch.meshgrid(*axes, indexing='ij'), -1)
        dense = self.density.get_dense_grid()
        alpha = F.max_pool3d(self.activate_density(dense.contiguous()), kernel_size=3, padding=1, stride=1)[0, 0]
        self.mask_cache = G.MaskGrid(path=None, mask=self.mask_cache(xyz) & (alpha > self.fast_color_thres),
                                     xyz_min=self.xyz_min, xyz_max=self.xyz_max).to(dev)

    @torch.no_grad()
    def update_occupancy_cache(self):
        """FourierGrid_model.py:441-456: mask_cache.mask &= max_pool3d(alpha(density(mask lattice))) > fast_color_thres, as two
        kernels (ops.lattice_alpha generates the lattice points in registers; ops.maxpool3_gt_and_ pools, thresholds and ANDs)."""
        _occupancy_update(self, self.density, self.mask_cache, float(self.voxel_size_ratio_density))

    @torch.no_grad()
    def maskout_near_cam_vox(self, cam_o, near_clip):
        """FourierGrid_model.py:375-388: density of the grid points closer than near_clip to any camera position (taken in each
        slab's own embedded coordinates gamma_i) is set to -100.  One kernel per slab: every voxel scans the camera list."""
        from . import ops
        dev = self.density.grid.device
        ind_norm = ((cam_o.to(dev) - self.xyz_min) / (self.xyz_max - self.xyz_min)).flip((-1,)) * 2 - 1
        F_ = self.density.nerf_pos_num_freq
        freqs = 2 ** torch.linspace(0, F_ - 1, F_, device=dev)
        emb = [ind_norm] + [f(fr * ind_norm) for fr in freqs for f in (torch.sin, torch.cos)]      # gamma_i of FourierGrid_grid.py:32-36
        for i, cam in enumerate(emb):
            # the reference writes `self.density.grid[0][i][...] = -100` (:388), which for its own [P,1,X,Y,Z] density grid raises an
            # IndexError at i = 1; the evident intent -- slab i, masked in slab i's embedded coordinates -- is grid[i][0]
            ops.maskout_near_cam_(self.density.grid.data[i][0], cam.reshape(-1, 3), near_clip, -100.0)

    def voxel_count_views(self, rays_o_tr, rays_d_tr, imsz, near, far, stepsize, downrate=1, irregular_shape=False):
        """FourierGrid_model.py:390-420: per-voxel number of training views that see it.  The reference materialises the sample
        points of 10 000 rays at a time and runs DenseGrid(ones).sum().backward(); here one kernel per view scatters the
        trilinear weights of every (ray, sample) straight into a per-view buffer, a second adds (buffer > 1) to the count."""
        from . import ops
        far = 1e9
        dev = self.density.grid.device
        ws = [int(v) for v in self.world_size_density]
        n_samples = int(np.linalg.norm(np.array(ws) + 1) / stepsize) + 1
        step = float(stepsize * self.voxel_size_density)
        mn, mx = self.density._bounds()
        count = torch.zeros(1, 1, *ws, device=dev)
        buf = torch.empty(*ws, device=dev)
        for rays_o_, rays_d_ in zip(rays_o_tr.split(imsz), rays_d_tr.split(imsz)):
            if not irregular_shape:
                rays_o_, rays_d_ = rays_o_[::downrate, ::downrate], rays_d_[::downrate, ::downrate]
            ro = rays_o_.to(dev).reshape(-1, 3).contiguous().float()
            rd = rays_d_.to(dev).reshape(-1, 3).contiguous().float()
            buf.zero_()
            ops.view_scatter_ones(ro, rd, mn, mx, ws, n_samples, near, far, step, buf)
            ops.count_gt_(count, buf, 1.0)
        return count

    def hit_coarse_geo(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """FourierGrid_model.py:495-507: does a ray hit the occupancy mask?"""
        from . import ops
        shape = rays_o.shape[:-1]
        rays_o = rays_o.reshape(-1, 3).contiguous()
        rays_d = rays_d.reshape(-1, 3).contiguous()
        ray_pts, mask_outbbox, ray_id = ops.sample_pts_on_rays(rays_o, rays_d, self.xyz_min, self.xyz_max, near, 1e9,
                                                               stepsize * float(self.voxel_size_density))[:3]
        mask_inbbox = ~mask_outbbox
        hit = torch.zeros([len(rays_o)], dtype=torch.bool, device=rays_o.device)
        hit[ray_id[mask_inbbox][self.mask_cache(ray_pts[mask_inbbox])]] = 1
        return hit.reshape(shape)

    def sample_ray(self, ori_rays_o, ori_rays_d, stepsize, is_train=False, **render_kwargs):
        """FourierGrid_model.py:509-552 return tuple (ray_pts, indexs, inner_mask, t, rays_d_extend)."""
        ray_pts, inner_mask, t = self._sample_dense(ori_rays_o, ori_rays_d, stepsize)
        return ray_pts, None, inner_mask, t, None

    def forward(self, rays_o, rays_d, viewdirs, global_step=None, is_train=False, **render_kwargs):
        assert len(rays_o.shape) == 2 and rays_o.shape[-1] == 3, 'Only support point queries in [N, 3] format'
        self._maybe_update_thres(global_step)
        if not self._fused_ok():
            return self.forward_ops(rays_o, rays_d, viewdirs, global_step=global_step, is_train=is_train, **render_kwargs)
        N = len(rays_o)
        (weights, alphainv_last, alpha, density, k0, ray_id, step_id, t, inner), t_table = self._march(
            rays_o, rays_d, render_kwargs['stepsize'])
        rgb = self._shade(k0, viewdirs, ray_id)
        rgb_marched = composite_rgb(weights, rgb, ray_id, N)
        if render_kwargs.get('rand_bkgd', False):
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_marched)
        s = 1 - 1 / (1 + t)
        ret = {'alphainv_last': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched, 'raw_density': density,
               'raw_alpha': alpha, 'raw_rgb': rgb, 'ray_id': ray_id, 'step_id': step_id, 'n_max': t_table.numel(),
               't': t, 's': s}
        if render_kwargs.get('render_depth', False):
            with torch.no_grad():
                ret['depth'] = segment_sum(weights * s, ray_id, N)
        return ret

    def forward_ops(self, rays_o, rays_d, viewdirs, global_step=None, is_train=False, **render_kwargs):
        """Op-by-op composition in the reference's order (FourierGrid_model.py:566-672)."""
        N = len(rays_o)
        dev = rays_o.device
        ray_pts, _, inner_mask, t, _ = self.sample_ray(rays_o, rays_d, **render_kwargs)
        n_max = len(t)
        S = n_max
        interval = render_kwargs['stepsize'] * self.voxel_size_ratio_density
        ray_id = torch.arange(N, device=dev).view(-1, 1).expand(N, S).flatten()
        step_id = torch.arange(S, device=dev).view(1, -1).expand(N, S).flatten()
        t = t[None].repeat(N, 1)
        density = self.density(ray_pts)
        alpha = self.activate_density(density, interval)
        if self.fast_color_thres > 0:
            mask = (alpha > self.fast_color_thres)
            ray_pts, inner_mask, t = ray_pts[mask], inner_mask[mask], t[mask]
            ray_id, step_id = ray_id[mask.flatten()], step_id[mask.flatten()]
            density, alpha = density[mask], alpha[mask]
        weights, alphainv_last = Alphas2Weights.apply(alpha.flatten().contiguous(), ray_id.contiguous(), N)
        if self.fast_color_thres > 0:
            mask = (weights > self.fast_color_thres)
            ray_pts, inner_mask, t = ray_pts[mask], inner_mask[mask], t[mask]
            ray_id, step_id = ray_id[mask], step_id[mask]
            density, alpha, weights = density[mask], alpha[mask], weights[mask]
        else:
            ray_pts = ray_pts.reshape(-1, 3)
            inner_mask = inner_mask.reshape(-1)
            t, density, alpha = t.reshape(-1), density.reshape(-1), alpha.reshape(-1)
        k0 = self.k0(ray_pts)
        rgb = self._shade(k0, viewdirs, ray_id)
        rgb_marched = composite_rgb(weights, rgb, ray_id, N)
        if render_kwargs.get('rand_bkgd', False):
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_marched)
        s = 1 - 1 / (1 + t)
        ret = {'alphainv_last': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched, 'raw_density': density,
               'raw_alpha': alpha, 'raw_rgb': rgb, 'ray_id': ray_id, 'step_id': step_id, 'n_max': n_max, 't': t, 's': s}
        if render_kwargs.get('render_depth', False):
            with torch.no_grad():
                ret['depth'] = segment_sum(weights * s, ray_id, N)
        return ret


# ======================================================================================================
class DirectContractedVoxGO(_ContractedBase):
    """FourierGrid/dcvgo.py:28-384 (DVGOv2 unbounded model): DenseGrid density/k0, cumdist_thres oversampling
    filter, mask cache, constant / random background term, wsum_mid."""

    T_BOUNDARY = 2.0
    USE_CUMDIST = True
    USE_MASKCACHE = True

    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, alpha_init=None, mask_cache_world_size=None,
                 fast_color_thres=0, bg_len=0.2, contracted_norm='inf', density_type='DenseGrid', k0_type='DenseGrid',
                 density_config={}, k0_config={}, rgbnet_dim=0, rgbnet_depth=3, rgbnet_width=128, viewbase_pe=4, **kwargs):
        super().__init__()
        self._init_scene(xyz_min, xyz_max, bg_len, fast_color_thres, contracted_norm)
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_base)
        self._set_grid_resolution(num_voxels)
        self.alpha_init = alpha_init
        self.register_buffer('act_shift', torch.FloatTensor([np.log(1 / (1 - alpha_init) - 1)]))
        self.density_type, self.density_config = density_type, density_config
        self.k0_type, self.k0_config = k0_type, k0_config
        if density_type != 'DenseGrid' or k0_type != 'DenseGrid':
            raise NotImplementedError('only DenseGrid is on the hot path (TensoRFGrid is out of scope, SURVEY.md 2 #6)')
        self.density = G.DenseGrid(channels=1, world_size=self.world_size, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        self.rgbnet_kwargs = dict(rgbnet_dim=rgbnet_dim, rgbnet_depth=rgbnet_depth, rgbnet_width=rgbnet_width,
                                  viewbase_pe=viewbase_pe)
        if rgbnet_dim <= 0:
            self.k0_dim = 3
            self.rgbnet = None
        else:
            self.k0_dim = rgbnet_dim
            self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
            self.rgbnet = _make_rgbnet(3 + 3 * viewbase_pe * 2 + rgbnet_dim, rgbnet_width, rgbnet_depth)
        self.k0 = G.DenseGrid(channels=self.k0_dim, world_size=self.world_size, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        if mask_cache_world_size is None:
            mask_cache_world_size = self.world_size
        self.mask_cache = G.MaskGrid(path=None, mask=torch.ones([int(v) for v in mask_cache_world_size], dtype=torch.bool),
                                     xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _set_grid_resolution(self, num_voxels):
        self.num_voxels = num_voxels
        self.voxel_size = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels)
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        self.world_len = self.world_size[0].item()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base

    def _world_len(self):
        return self.world_len

    def _voxel_size_ratio(self):
        return self.voxel_size_ratio

    def get_kwargs(self):
        return {
            'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(),
            'num_voxels': self.num_voxels, 'num_voxels_base': self.num_voxels_base, 'alpha_init': self.alpha_init,
            'voxel_size_ratio': self.voxel_size_ratio, 'mask_cache_world_size': list(self.mask_cache.mask.shape),
            'fast_color_thres': self.fast_color_thres, 'contracted_norm': self.contracted_norm,
            'density_type': self.density_type, 'k0_type': self.k0_type, 'density_config': self.density_config,
            'k0_config': self.k0_config, **self.rgbnet_kwargs,
        }

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels):
        self._set_grid_resolution(num_voxels)
        self.density.scale_volume_grid(self.world_size)
        self.k0.scale_volume_grid(self.world_size)
        if np.prod(self.world_size.tolist()) <= 256 ** 3:
            dev = self.density.grid.device
            ws = [int(v) for v in self.world_size]
            axes = [torch.linspace(float(self.xyz_min[a]), float(self.xyz_max[a]), ws[a], device=dev) for a in range(3)]
            xyz = torch.stack(torch.meshgrid(*axes, indexing='ij'), -1)
            alpha = F.max_pool3d(self.activate_density(self.density.get_dense_grid().contiguous()), kernel_size=3,
                                 padding=1, stride=1)[0, 0]
            self.mask_cache = G.MaskGrid(path=None, mask=self.mask_cache(xyz) & (alpha > self.fast_color_thres),
                                         xyz_min=self.xyz_min, xyz_max=self.xyz_max).to(dev)
            self._mask_host = None

    @torch.no_grad()
    def update_occupancy_cache(self):
        """dcvgo.py:214-226 (same composition as FourierGrid_model.py:441-456)."""
        _occupancy_update(self, self.density, self.mask_cache, float(self.voxel_size_ratio))
        self._mask_host = None

    def sample_ray(self, ori_rays_o, ori_rays_d, stepsize, is_train=False, **render_kwargs):
        """dcvgo.py:228-262 return tuple (ray_pts, inner_mask, t)."""
        return self._sample_dense(ori_rays_o, ori_rays_d, stepsize)

    def _finish(self, N, dev, weights, alphainv_last, density, alpha, rgb, ray_id, step_id, t, inner_mask, n_max,
                is_train, render_kwargs):
        rgb_marched = composite_rgb(weights, rgb, ray_id, N)
        if render_kwargs.get('rand_bkgd', False) and is_train:
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * torch.rand_like(rgb_marched)
        else:
            rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * render_kwargs['bg']
        wsum_mid = segment_sum(weights[inner_mask], ray_id[inner_mask], N)
        s = 1 - 1 / (1 + t)
        ret = {'alphainv_last': alphainv_last, 'weights': weights, 'wsum_mid': wsum_mid, 'rgb_marched': rgb_marched,
               'raw_density': density, 'raw_alpha': alpha, 'raw_rgb': rgb, 'ray_id': ray_id, 'step_id': step_id,
               'n_max': n_max, 't': t, 's': s}
        if render_kwargs.get('render_depth', False):
            with torch.no_grad():
                ret['depth'] = segment_sum(weights * s, ray_id, N)
        return ret

    def forward(self, rays_o, rays_d, viewdirs, global_step=None, is_train=False, **render_kwargs):
        assert len(rays_o.shape) == 2 and rays_o.shape[-1] == 3, 'Only support point queries in [N, 3] format'
        self._maybe_update_thres(global_step)
        if not self._fused_ok():
            return self.forward_ops(rays_o, rays_d, viewdirs, global_step=global_step, is_train=is_train, **render_kwargs)
        N = len(rays_o)
        # render_kwargs['coherent_rays'] (set by render.render_rays: a frame's image-ordered chunks) selects the TMA-staged
        # feature read; it is a hint, never a requirement -- any rays give the same result
        (weights, alphainv_last, alpha, density, k0, ray_id, step_id, t, inner), t_table = self._march(
            rays_o, rays_d, render_kwargs['stepsize'], coherent=bool(render_kwargs.get('coherent_rays', False)))
        rgb = self._shade(k0, viewdirs, ray_id)
        return self._finish(N, rays_o.device, weights, alphainv_last, density, alpha, rgb, ray_id, step_id, t, inner,
                            t_table.numel(), is_train, render_kwargs)

    def forward_ops(self, rays_o, rays_d, viewdirs, global_step=None, is_train=False, **render_kwargs):
        """Op-by-op composition in the reference's order (dcvgo.py:275-384)."""
        from . import ops
        N = len(rays_o)
        dev = rays_o.device
        ray_pts, inner_mask, t = self.sample_ray(rays_o, rays_d, is_train=global_step is not None, **render_kwargs)
        n_max = len(t)
        S = n_max
        interval = render_kwargs['stepsize'] * self.voxel_size_ratio
        ray_id = torch.arange(N, device=dev).view(-1, 1).expand(N, S).flatten()
        step_id = torch.arange(S, device=dev).view(1, -1).expand(N, S).flatten()
        mask = inner_mask.clone()
        dist_thres = (2 + 2 * self.bg_len) / self.world_len * render_kwargs['stepsize'] * 0.95
        dist = (ray_pts[:, 1:] - ray_pts[:, :-1]).norm(dim=-1)
        mask[:, 1:] |= ops.cumdist_thres(dist.contiguous(), dist_thres)
        ray_pts, inner_mask = ray_pts[mask], inner_mask[mask]
        t = t[None].repeat(N, 1)[mask]
        ray_id, step_id = ray_id[mask.flatten()], step_id[mask.flatten()]
        mask = self.mask_cache(ray_pts)
        ray_pts, inner_mask, t, ray_id, step_id = ray_pts[mask], inner_mask[mask], t[mask], ray_id[mask], step_id[mask]
        density = self.density(ray_pts)
        alpha = self.activate_density(density, interval)
        if self.fast_color_thres > 0:
            mask = (alpha > self.fast_color_thres)
            ray_pts, inner_mask, t, ray_id, step_id = ray_pts[mask], inner_mask[mask], t[mask], ray_id[mask], step_id[mask]
            density, alpha = density[mask], alpha[mask]
        weights, alphainv_last = Alphas2Weights.apply(alpha.contiguous(), ray_id.contiguous(), N)
        if self.fast_color_thres > 0:
            mask = (weights > self.fast_color_thres)
            ray_pts, inner_mask, t, ray_id, step_id = ray_pts[mask], inner_mask[mask], t[mask], ray_id[mask], step_id[mask]
            density, alpha, weights = density[mask], alpha[mask], weights[mask]
        k0 = self.k0(ray_pts)
        rgb = self._shade(k0, viewdirs, ray_id)
        return self._finish(N, dev, weights, alphainv_last, density, alpha, rgb, ray_id, step_id, t, inner_mask, n_max,
                            is_train, render_kwargs)


# ======================================================================================================
class DirectVoxGO(nn.Module):
    """Bounded-scene model (FourierGrid/dvgo.py:26-425): ragged AABB sampling (sample_pts_on_rays) instead of the
    contracted schedule; depth = sum w * step_id.  Composed from the drop-in ops (BASELINE config 1 family)."""

    def __init__(self, xyz_min, xyz_max, num_voxels=0, num_voxels_base=0, alpha_init=None, mask_cache_path=None,
                 mask_cache_thres=1e-3, mask_cache_world_size=None, fast_color_thres=0, density_type='DenseGrid',
                 k0_type='DenseGrid', density_config={}, k0_config={}, rgbnet_dim=0, rgbnet_direct=False,
                 rgbnet_full_implicit=False, rgbnet_depth=3, rgbnet_width=128, viewbase_pe=4, **kwargs):
        super().__init__()
        if rgbnet_full_implicit:
            raise NotImplementedError('rgbnet_full_implicit is outside the hot-path scope')
        self.register_buffer('xyz_min', torch.as_tensor(np.asarray(xyz_min), dtype=torch.float32))
        self.register_buffer('xyz_max', torch.as_tensor(np.asarray(xyz_max), dtype=torch.float32))
        self.fast_color_thres = fast_color_thres
        self.num_voxels_base = num_voxels_base
        self.voxel_size_base = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels_base)
        self.alpha_init = alpha_init
        self.register_buffer('act_shift', torch.FloatTensor([np.log(1 / (1 - alpha_init) - 1)]))
        self._set_grid_resolution(num_voxels)
        self.density_type, self.density_config, self.k0_type, self.k0_config = density_type, density_config, k0_type, k0_config
        self.density = G.DenseGrid(channels=1, world_size=self.world_size, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        self.rgbnet_kwargs = dict(rgbnet_dim=rgbnet_dim, rgbnet_direct=rgbnet_direct, rgbnet_full_implicit=rgbnet_full_implicit,
                                  rgbnet_depth=rgbnet_depth, rgbnet_width=rgbnet_width, viewbase_pe=viewbase_pe)
        self.rgbnet_direct = rgbnet_direct
        if rgbnet_dim <= 0:
            self.k0_dim, self.rgbnet = 3, None
        else:
            self.k0_dim = rgbnet_dim
            self.register_buffer('viewfreq', torch.FloatTensor([(2 ** i) for i in range(viewbase_pe)]))
            dim0 = 3 + 3 * viewbase_pe * 2 + (rgbnet_dim if rgbnet_direct else rgbnet_dim - 3)
            self.rgbnet = _make_rgbnet(dim0, rgbnet_width, rgbnet_depth)
        self.k0 = G.DenseGrid(channels=self.k0_dim, world_size=self.world_size, xyz_min=self.xyz_min, xyz_max=self.xyz_max)
        self.mask_cache_path, self.mask_cache_thres = mask_cache_path, mask_cache_thres
        if mask_cache_world_size is None:
            mask_cache_world_size = self.world_size
        if mask_cache_path:
            raise NotImplementedError('coarse-checkpoint mask cache needs a device at construction; build MaskGrid(path=...) '
                                      'and assign model.mask_cache instead')
        self.mask_cache = G.MaskGrid(path=None, mask=torch.ones([int(v) for v in mask_cache_world_size], dtype=torch.bool),
                                     xyz_min=self.xyz_min, xyz_max=self.xyz_max)

    def _set_grid_resolution(self, num_voxels):
        self.num_voxels = num_voxels
        self.voxel_size = _cube_root_size(self.xyz_min, self.xyz_max, num_voxels)
        self.world_size = ((self.xyz_max - self.xyz_min) / self.voxel_size).long()
        self.voxel_size_ratio = self.voxel_size / self.voxel_size_base

    def get_kwargs(self):
        return {
            'xyz_min': self.xyz_min.cpu().numpy(), 'xyz_max': self.xyz_max.cpu().numpy(), 'num_voxels': self.num_voxels,
            'num_voxels_base': self.num_voxels_base, 'alpha_init': self.alpha_init, 'voxel_size_ratio': self.voxel_size_ratio,
            'mask_cache_path': self.mask_cache_path, 'mask_cache_thres': self.mask_cache_thres,
            'mask_cache_world_size': list(self.mask_cache.mask.shape), 'fast_color_thres': self.fast_color_thres,
            'density_type': self.density_type, 'k0_type': self.k0_type, 'density_config': self.density_config,
            'k0_config': self.k0_config, **self.rgbnet_kwargs,
        }

    def activate_density(self, density, interval=None):
        interval = interval if interval is not None else self.voxel_size_ratio
        shape = density.shape
        return Raw2Alpha.apply(density.flatten().contiguous(), self.act_shift, interval).reshape(shape)

    def density_total_variation_add_grad(self, weight, dense_mode):
        w = weight * float(self.world_size.max()) / 128
        self.density.total_variation_add_grad(w, w, w, dense_mode)

    def k0_total_variation_add_grad(self, weight, dense_mode):
        w = weight * float(self.world_size.max()) / 128
        self.k0.total_variation_add_grad(w, w, w, dense_mode)

    def tv_terms(self, weight_density=0., weight_k0=0., dense_mode=True):
        """{grid parameter: (wx, wy, wz, dense_mode)} for dist.reduce_tv_step (same weights as the two methods above)."""
        out = {}
        for grid, weight in ((self.density, weight_density), (self.k0, weight_k0)):
            if weight > 0:
                w = weight * float(self.world_size.max()) / 128
                out[grid.grid] = (w, w, w, dense_mode)
        return out

    def sample_ray(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """dvgo.py:306-328."""
        from . import ops
        far = 1e9
        stepdist = stepsize * float(self.voxel_size)
        ray_pts, mask_outbbox, ray_id, step_id, N_steps, t_min, t_max = ops.sample_pts_on_rays(
            rays_o.contiguous(), rays_d.contiguous(), self.xyz_min, self.xyz_max, near, far, stepdist)
        mask_inbbox = ~mask_outbbox
        return ray_pts[mask_inbbox], ray_id[mask_inbbox], step_id[mask_inbbox]

    def hit_coarse_geo(self, rays_o, rays_d, near, far, stepsize, **render_kwargs):
        """dvgo.py:292-304."""
        from . import ops
        shape = rays_o.shape[:-1]
        rays_o = rays_o.reshape(-1, 3).contiguous()
        rays_d = rays_d.reshape(-1, 3).contiguous()
        ray_pts, mask_outbbox, ray_id = ops.sample_pts_on_rays(rays_o, rays_d, self.xyz_min, self.xyz_max, near, 1e9,
                                                               stepsize * float(self.voxel_size))[:3]
        mask_inbbox = ~mask_outbbox
        hit = torch.zeros([len(rays_o)], dtype=torch.bool, device=rays_o.device)
        hit[ray_id[mask_inbbox][self.mask_cache(ray_pts[mask_inbbox])]] = 1
        return hit.reshape(shape)

    def forward(self, rays_o, rays_d, viewdirs, global_step=None, **render_kwargs):
        assert len(rays_o.shape) == 2 and rays_o.shape[-1] == 3, 'Only suuport point queries in [N, 3] format'
        N, dev = len(rays_o), rays_o.device
        ray_pts, ray_id, step_id = self.sample_ray(rays_o=rays_o, rays_d=rays_d, **render_kwargs)
        interval = render_kwargs['stepsize'] * self.voxel_size_ratio
        if self.mask_cache is not None:
            mask = self.mask_cache(ray_pts)
            ray_pts, ray_id, step_id = ray_pts[mask], ray_id[mask], step_id[mask]
        density = self.density(ray_pts)
        alpha = self.activate_density(density, interval)
        if self.fast_color_thres > 0:
            mask = (alpha > self.fast_color_thres)
            ray_pts, ray_id, step_id, density, alpha = ray_pts[mask], ray_id[mask], step_id[mask], density[mask], alpha[mask]
        weights, alphainv_last = Alphas2Weights.apply(alpha.contiguous(), ray_id.contiguous(), N)
        if self.fast_color_thres > 0:
            mask = (weights > self.fast_color_thres)
            weights, alpha, ray_pts, ray_id, step_id = weights[mask], alpha[mask], ray_pts[mask], ray_id[mask], step_id[mask]
        k0 = self.k0(ray_pts)
        if self.rgbnet is None:
            rgb = torch.sigmoid(k0)
        else:
            k0_view = k0 if self.rgbnet_direct else k0[:, 3:]
            emb = _view_embed(viewdirs, self.viewfreq).flatten(0, -2)[ray_id]
            logit = self.rgbnet(torch.cat([k0_view, emb], -1))
            rgb = torch.sigmoid(logit if self.rgbnet_direct else logit + k0[:, :3])
        rgb_marched = composite_rgb(weights, rgb, ray_id, N)
        rgb_marched = rgb_marched + alphainv_last.unsqueeze(-1) * render_kwargs['bg']
        ret = {'alphainv_last': alphainv_last, 'weights': weights, 'rgb_marched': rgb_marched, 'raw_alpha': alpha,
               'raw_rgb': rgb, 'ray_id': ray_id}
        if render_kwargs.get('render_depth', False):
            with torch.no_grad():
                ret['depth'] = segment_sum(weights * step_id, ray_id, N)
        return ret
