"""Camera rays and training batches with the reference's names and signatures (FourierGrid/dvgo.py:490-667), built on
device: one launch per view for rays_o / rays_d / viewdirs (``ubn_get_rays_of_a_view``) and one launch per step for the
batch gather (``ubn_gather_rays``) instead of ~20 torch kernels per view and four index kernels per step.

Drop-in for ``dvgo.get_rays / ndc_rays / get_rays_of_a_view / get_training_rays / get_training_rays_flatten /
get_training_rays_in_maskcache_sampling / batch_indices_generator``; the extra ``gather_ray_batch`` is the fused form of
run_train.py:204-212.  Outputs live on the device of ``c2w`` when that is a CUDA tensor, else on the current CUDA device
(the reference builds rays on ``c2w.device`` and copies them to the image device; with images on the GPU this is the same).
"""
import ctypes

import numpy as np
import torch

from ._cabi import check, ptr, stream_of
from .ops import _Guard

_MODES = {'lefttop': 0, 'center': 1, 'random': 2}


def _host_f32(x, n):
    a = np.ascontiguousarray(np.asarray(x.detach().cpu() if torch.is_tensor(x) else x, dtype=np.float32))
    if a.size < n:
        raise RuntimeError('camera matrix too small')
    return a


def _device_of(c2w, device=None):
    if device is not None:
        return torch.device(device)
    if torch.is_tensor(c2w) and c2w.is_cuda:
        return c2w.device
    return torch.device('cuda', torch.cuda.current_device())


def _rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode, device=None, out=None):
    if mode not in _MODES:
        raise NotImplementedError
    H, W = int(H), int(W)
    dev = _device_of(c2w, device)
    K_h = _host_f32(K, 9).reshape(-1)
    c_h = _host_f32(c2w, 12)
    stride = c_h.shape[-1]
    c_h = c_h.reshape(-1)
    if out is None:
        out = [torch.empty(H, W, 3, device=dev) for _ in range(3)]
    else:                                   # caller-provided contiguous fp32 CUDA buffers of H*W*3 elements (flattened views)
        assert all(o.is_cuda and o.is_contiguous() and o.dtype == torch.float32 and o.numel() == H * W * 3 for o in out)
        dev = out[0].device
    jitter = torch.rand(2, H, W, device=dev) if mode == 'random' else None
    with _Guard(out[0]) as lib:
        check(lib.ubn_get_rays_of_a_view(H, W, K_h.ctypes.data_as(ctypes.c_void_p), c_h.ctypes.data_as(ctypes.c_void_p),
                                         int(stride), int(bool(ndc)), int(bool(inverse_y)), int(bool(flip_x)),
                                         int(bool(flip_y)), _MODES[mode], ptr(jitter), ptr(out[0]), ptr(out[1]), ptr(out[2]),
                                         stream_of(out[0])))
    return out


def get_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode='center'):
    """dvgo.py:552-557 -> (rays_o, rays_d, viewdirs), each [H,W,3]."""
    return tuple(_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, mode))


def get_rays(H, W, K, c2w, inverse_y, flip_x, flip_y, mode='center'):
    """dvgo.py:492-520 -> (rays_o, rays_d)."""
    o, d, _ = _rays_of_a_view(H, W, K, c2w, False, inverse_y, flip_x, flip_y, mode)
    return o, d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """dvgo.py:532-550 (elementwise torch; the fused path is get_rays_of_a_view(ndc=True))."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (W / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (H / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


@torch.no_grad()
def get_training_rays(rgb_tr, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """dvgo.py:560-582: all views share (H, W, K) -> [n_views,H,W,3] ray tensors on rgb_tr's device."""
    assert len(np.unique(HW, axis=0)) == 1
    assert len(np.unique(np.asarray(Ks).reshape(len(Ks), -1), axis=0)) == 1
    assert len(rgb_tr) == len(train_poses) and len(rgb_tr) == len(Ks) and len(rgb_tr) == len(HW)
    H, W = HW[0]
    K = Ks[0]
    dev = rgb_tr.device if rgb_tr.is_cuda else None
    per_view = [_rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, 'center', dev) for c2w in train_poses]
    rays_o_tr, rays_d_tr, viewdirs_tr = (torch.stack([v[k] for v in per_view]).to(rgb_tr.device) for k in range(3))
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, [1] * len(rgb_tr)


@torch.no_grad()
def get_training_rays_flatten(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y):
    """dvgo.py:585-612: views of different sizes flattened to [N,3]; the per-view kernels write straight into the slices."""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    dev_img = rgb_tr_ori[0].device
    dev = dev_img if dev_img.type == 'cuda' else None
    N = sum(im.shape[0] * im.shape[1] for im in rgb_tr_ori)
    rgb_tr = torch.zeros([N, 3], device=dev_img)
    gpu = dev if dev is not None else torch.device('cuda', torch.cuda.current_device())
    rays = [torch.empty(N, 3, device=gpu) for _ in range(3)]
    imsz, top = [], 0
    for c2w, img, (H, W), K in zip(train_poses, rgb_tr_ori, HW, Ks):
        assert img.shape[:2] == (H, W)
        n = H * W
        _rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, 'center', out=[r[top:top + n] for r in rays])
        rgb_tr[top:top + n].copy_(img.flatten(0, 1))
        imsz.append(n)
        top += n
    assert top == N
    rays_o_tr, rays_d_tr, viewdirs_tr = (r.to(dev_img) for r in rays)
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


@torch.no_grad()
def get_training_rays_in_maskcache_sampling(rgb_tr_ori, train_poses, HW, Ks, ndc, inverse_y, flip_x, flip_y, model,
                                            render_kwargs):
    """dvgo.py:615-655: keep only rays that hit the coarse geometry (model.hit_coarse_geo)."""
    assert len(rgb_tr_ori) == len(train_poses) and len(rgb_tr_ori) == len(Ks) and len(rgb_tr_ori) == len(HW)
    CHUNK = 64
    dev_img = rgb_tr_ori[0].device
    dev = dev_img if dev_img.type == 'cuda' else None
    keep = [[], [], [], []]
    imsz = []
    for c2w, img, (H, W), K in zip(train_poses, rgb_tr_ori, HW, Ks):
        assert img.shape[:2] == (H, W)
        rays_o, rays_d, viewdirs = _rays_of_a_view(H, W, K, c2w, ndc, inverse_y, flip_x, flip_y, 'center', dev)
        mask = torch.empty(img.shape[:2], device=rays_o.device, dtype=torch.bool)
        for i in range(0, img.shape[0], CHUNK):
            mask[i:i + CHUNK] = model.hit_coarse_geo(rays_o=rays_o[i:i + CHUNK], rays_d=rays_d[i:i + CHUNK], **render_kwargs)
        for lst, t in zip(keep, (img.to(rays_o.device), rays_o, rays_d, viewdirs)):
            lst.append(t[mask].to(dev_img))
        imsz.append(int(mask.sum()))
    rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr = (torch.cat(lst) for lst in keep)
    return rgb_tr, rays_o_tr, rays_d_tr, viewdirs_tr, imsz


def batch_indices_generator(N, BS):
    """dvgo.py:658-667 (host permutation like the reference: np.random.permutation seeds stay comparable)."""
    idx, top = torch.LongTensor(np.random.permutation(N)), 0
    while True:
        if top + BS > N:
            idx, top = torch.LongTensor(np.random.permutation(N)), 0
        yield idx[top:top + BS]
        top += BS


def gather_ray_batch(sel_i, *arrays, validate_device_indices=True):
    """run_train.py:204-212 in one launch: ``[a[sel_i] for a in arrays]`` for up to four [N,3] fp32 CUDA arrays."""
    if not 1 <= len(arrays) <= 4:
        raise ValueError('gather_ray_batch takes 1..4 arrays')
    a0 = arrays[0]
    for a in arrays:
        if not (a.is_cuda and a.dtype == torch.float32 and a.dim() == 2 and a.shape[1] == 3 and a.is_contiguous()
                and a.shape[0] == a0.shape[0]):
            raise RuntimeError('arrays must be contiguous CUDA fp32 [N,3] tensors of the same length')
    if not sel_i.is_cuda:
        # validate where the indices already live (batch_indices_generator yields CPU LongTensors): no device round trip
        if sel_i.numel() and (int(sel_i.min()) < -a0.shape[0] or int(sel_i.max()) >= a0.shape[0]):   # negatives wrap like Python
            raise IndexError('index out of range in gather_ray_batch')
        check_device = False
    else:
        check_device = validate_device_indices
    sel = sel_i.to(device=a0.device, dtype=torch.int64, non_blocking=True).contiguous()
    n = sel.numel()
    outs = [torch.empty(n, 3, device=a0.device) for _ in arrays]
    if n == 0:
        return outs
    oob = torch.zeros(1, dtype=torch.int32, device=a0.device)
    src = (ctypes.c_void_p * len(arrays))(*[a.data_ptr() for a in arrays])
    dst = (ctypes.c_void_p * len(arrays))(*[o.data_ptr() for o in outs])
    with _Guard(a0) as lib:
        check(lib.ubn_gather_rays(src, dst, len(arrays), ptr(sel), n, a0.shape[0], ptr(oob), stream_of(a0)))
    # the kernel never reads out of range (an invalid index is skipped and raises the flag); reading the flag is a blocking
    # host sync, so it is only done for device-resident indices and only when asked for
    if check_device and int(oob.item()):
        raise IndexError('index out of range in gather_ray_batch')
    return outs
