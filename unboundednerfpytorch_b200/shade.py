"""rgbnet (feature -> RGB MLP) as one fused kernel per direction (csrc/shade.cu) behind an autograd.Function.

rgb = sigmoid(rgbnet(cat[k0, viewdirs_emb[ray_id]]))   (FourierGrid_model.py:631-637, dcvgo.py:337-342)

The view-direction half of the first Linear is constant along a ray, so the host folds it into a per-ray bias table
``vb = emb(viewdirs) @ W1[:, 12:].T + b1`` ([N,128], one tiny GEMM that torch differentiates for dW1[:,12:], db1) and the
kernel runs the per-sample part: 12 -> 128 -> 128 -> 3 with the activations resident on chip, fp32 arithmetic.
"""
import torch

import os

from . import _cabi, ops
from ._cabi import c_i64, c_int, check, ptr, stream_of

# engine: 'tc3' = tcgen05 3xTF32 (fp32-grade, the default and the only mode the 1e-5 parity tests accept), 'tc1' = tcgen05 with a
# single TF32 pass per product in the forward AND (with BWD_MODE 'fused') the backward -- the opt-in reduced-precision training
# mode, ~1e-3 relative error, gated by the PSNR test in tests/test_gpu_models.py --, 'simt' = fp32 FFMA
MODE = os.environ.get('UBN_RGBNET_MODE', 'tc3')
# backward engine: 'fused' = tcgen05 3xTF32, dZ2 -> dH1 -> dZ1 -> dX chained through tensor memory + all sample reductions in one
# warp-specialised kernel (4 row warps drive the tensor cores, 4 column warps reduce over samples), dW2 in a second launch; no
# intermediate in HBM; 'fused4' = the same without warp specialisation (A/B); 'tc3' = the previous three-launch form (dZ1 round trip + CUDA-core kernel for
# the small gradients; kept for A/B); 'simt' = fp32 FFMA
BWD_MODE = os.environ.get('UBN_RGBNET_BWD_MODE', 'fused')
# ReLU masks instead of activation re-reads in the fused backward (default on): the forward leaves the masks of H1 (16 B per sample)
# so that launch 1 gates dH1 without loading the H1 rows; launch 1 ballots the masks of H2 so that the dW2 launch rebuilds dZ2 (and
# sums db2) without reading H2 a second time.  Off = both launches re-read the saves (A/B; tests cover both).
USE_MASKS = os.environ.get('UBN_RGBNET_MASKS', '1') == '1'


class _ShadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, vb, ray_id, W1k, W2, b2, W3, b3, need_grad):
        feat, vb, ray_id = feat.contiguous(), vb.contiguous(), ray_id.contiguous()
        W1k, W2, b2, W3, b3 = (t.contiguous() for t in (W1k, W2, b2, W3, b3))
        M = feat.shape[0]
        dev = feat.device
        rgb = torch.empty(M, 3, dtype=torch.float32, device=dev)
        # need_grad comes from the caller (shade()): inside Function.forward grad mode is always off, and needs_input_grad
        # mirrors requires_grad of the inputs even under torch.no_grad() -- render / eval forwards must not allocate and stream
        # the two [M,128] activation saves
        need_grad = bool(need_grad) and any(ctx.needs_input_grad)
        # panel-layout saves ([tile][32 column quads][128 rows][4], ceil(M/128)*128 rows): coalesced for the row-per-thread kernels
        # on both sides; only the tcgen05 forward writes it and only the warp-specialised backward (+ dW2) reads it
        panel = need_grad and MODE in ('tc3', 'tc1', 'tc3w4') and BWD_MODE == 'fused'
        rows = -(-M // 128) * 128 if panel else M
        h1 = torch.empty(rows, 128, dtype=torch.float32, device=dev) if need_grad else None
        h2 = torch.empty(rows, 128, dtype=torch.float32, device=dev) if need_grad else None
        # ReLU masks of H1 (16 B per sample): the first backward launch gates dH1 with them instead of loading the 512-byte H1 rows
        m1 = torch.empty(rows * 4, dtype=torch.int32, device=dev) if (panel and USE_MASKS) else None
        with ops._Guard(feat) as lib:
            with _cabi.timed('rgbnet_fwd'):
                if MODE in ('tc3', 'tc1', 'tc3w4'):      # 'tc3w4': the 4-warp form of the forward kernel (A/B of the 8-warp default)
                    check(lib.ubn_rgbnet_fwd_tc(ptr(feat), ptr(vb), ptr(ray_id), ptr(W1k), ptr(W2), ptr(b2), ptr(W3), ptr(b3),
                                                c_i64(M), ptr(rgb), ptr(h1), ptr(h2), ptr(m1),
                                                c_int((1 if MODE == 'tc1' else 0) | (2 if MODE == 'tc3w4' else 0) | (4 if panel else 0)),
                                                stream_of(feat)))
                else:
                    check(lib.ubn_rgbnet_fwd(ptr(feat), ptr(vb), ptr(ray_id), ptr(W1k), ptr(W2), ptr(b2), ptr(W3), ptr(b3),
                                             c_i64(M), ptr(rgb), ptr(h1), ptr(h2), stream_of(feat)))
        if need_grad:
            ctx.save_for_backward(feat, ray_id, W1k, W2, W3, rgb, h1, h2)
            ctx.m1 = m1
            ctx.n_rays = vb.shape[0]
            ctx.panel = panel
            ctx.bwd_mode = BWD_MODE
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb):
        feat, ray_id, W1k, W2, W3, rgb, h1, h2 = ctx.saved_tensors
        dev = feat.device
        M = feat.shape[0]
        g_rgb = g_rgb.contiguous()
        g_feat = torch.empty_like(feat)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        g_vb, gW1k, gW2, gb2, gW3, gb3 = z(ctx.n_rays, 128), z(128, 12), z(128, 128), z(128), z(3, 128), z(3)
        with ops._Guard(feat) as lib:
            bwd_mode = ctx.bwd_mode                      # as chosen in forward (the save layout depends on it)
            if bwd_mode in ('fused', 'fused4'):          # 'fused4': the same kernel without warp specialisation (A/B)
                # panel saves: launch 1 leaves the ReLU masks of H2 (2 KB per 128-sample tile) in this scratch and the dW2 launch
                # rebuilds dZ2 from them instead of reading the 512 B/sample of H2 again
                masks = torch.empty(-(-M // 128) * 512, dtype=torch.int32, device=dev) if (ctx.panel and ctx.m1 is not None) else None
                with _cabi.timed('rgbnet_bwd'):
                    check(lib.ubn_rgbnet_bwd_tc_fused(ptr(feat), ptr(ray_id), ptr(W1k), ptr(W2), ptr(W3), ptr(rgb), ptr(h1), ptr(h2),
                                                      ptr(g_rgb), c_i64(M), ptr(g_feat), ptr(g_vb), ptr(gW1k), ptr(gW2), ptr(gb2),
                                                      ptr(gW3), ptr(gb3), ptr(masks), ptr(ctx.m1 if bwd_mode == 'fused' else None), c_int((1 if MODE == 'tc1' else 0) | (2 if bwd_mode == 'fused4' else 0) | (4 if ctx.panel else 0)),
                                                      stream_of(feat)))
                return g_feat, g_vb, None, gW1k, gW2, gb2, gW3, gb3, None
            if bwd_mode == 'tc3':
                dz1 = torch.empty(M, 128, dtype=torch.float32, device=dev)
                with _cabi.timed('rgbnet_bwd'):
                    check(lib.ubn_rgbnet_bwd_tc_data(ptr(W2), ptr(W3), ptr(rgb), ptr(h1), ptr(h2), ptr(g_rgb), c_i64(M),
                                                     ptr(dz1), ptr(gW2), stream_of(feat)))
                with _cabi.timed('rgbnet_bwd_small'):
                    check(lib.ubn_rgbnet_bwd_small(ptr(feat), ptr(ray_id), ptr(W1k), ptr(W3), ptr(rgb), ptr(h2), ptr(g_rgb),
                                                   ptr(dz1), c_i64(M), ptr(g_feat), ptr(g_vb), ptr(gW1k), ptr(gb2), ptr(gW3),
                                                   ptr(gb3), stream_of(feat)))
                return g_feat, g_vb, None, gW1k, gW2, gb2, gW3, gb3, None
            with _cabi.timed('rgbnet_bwd'):
                check(lib.ubn_rgbnet_bwd(ptr(feat), ptr(ray_id), ptr(W1k), ptr(W2), ptr(W3), ptr(rgb), ptr(h1), ptr(h2),
                                         ptr(g_rgb), c_i64(M), ptr(g_feat), ptr(g_vb), ptr(gW1k), ptr(gW2), ptr(gb2),
                                         ptr(gW3), ptr(gb3), stream_of(feat)))
        return g_feat, g_vb, None, gW1k, gW2, gb2, gW3, gb3, None


def supported(rgbnet, k0_dim):
    """3-layer, width-128 rgbnet on 12 features (rgbnet_depth=3, rgbnet_width=128, rgbnet_dim=12: every shipped config)."""
    try:
        l1, l2, l3 = rgbnet[0], rgbnet[2][0], rgbnet[3]
    except Exception:
        return False
    return (len(rgbnet) == 4 and k0_dim == 12 and l1.weight.shape[0] == 128 and tuple(l2.weight.shape) == (128, 128)
            and tuple(l3.weight.shape) == (3, 128) and l1.weight.is_cuda and l1.weight.dtype == torch.float32)


def shade(rgbnet, k0, view_emb, ray_id):
    """k0 [M,12], view_emb [N,27] (cat[v, sin, cos]), ray_id [M] sorted -> rgb [M,3]."""
    l1, l2, l3 = rgbnet[0], rgbnet[2][0], rgbnet[3]
    kd = k0.shape[1]
    vb = torch.addmm(l1.bias, view_emb, l1.weight[:, kd:].t())
    return _ShadeFn.apply(k0, vb, ray_id, l1.weight[:, :kd], l2.weight, l2.bias, l3.weight, l3.bias, torch.is_grad_enabled())
