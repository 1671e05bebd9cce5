"""Drop-in `inverse_warp` module (reference: inverse_warp.py): `inverse_warp(img, depth, pose, intrinsics,
intrinsics_inv, rotation_mode='euler', padding_mode='zeros')` -> projected image, executed by one fused CUDA
kernel (pixel2cam -> pose_vec2mat -> cam2pixel -> bilinear grid_sample; reference :26-193) with an analytic
fused backward (depth, pose and optionally image gradients).

`align_corners`: the reference calls `F.grid_sample` without the flag (:191); under the torch it is executed
with today that means align_corners=False, which is the default here (SURVEY.md hard part 6).  Pass
align_corners=True for torch<=1.2 behaviour.
"""
import torch

from . import _lib as L


def check_sizes(input, input_name, expected):
    condition = [input.ndimension() == len(expected)]
    for i, size in enumerate(expected):
        if size.isdigit():
            condition.append(input.size(i) == int(size))
    assert all(condition), "wrong size for {}, expected {}, got  {}".format(input_name, 'x'.join(expected),
                                                                              list(input.size()))


_ROT = {'euler': 0, 'quat': 1}
_PAD = {'zeros': 0, 'border': 1}


class _InverseWarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, pose, K, Kinv, rot, pad, align):
        L.require_cuda(img, depth, pose, K, Kinv)
        img, depth, pose = img.contiguous().float(), depth.contiguous().float(), pose.contiguous().float()
        K, Kinv = K.contiguous().float(), Kinv.contiguous().float()
        B, Cc, h, w = img.shape
        out = torch.empty_like(img)
        L.call('dn_inverse_warp_fwd', L.ptr(img), L.ptr(depth), L.ptr(pose), L.ptr(K), L.ptr(Kinv), B, Cc, h, w, rot, pad,
               align, L.ptr(out), L.stream_ptr())
        ctx.save_for_backward(img, depth, pose, K, Kinv)
        ctx.cfg = (rot, pad, align)
        return out

    @staticmethod
    def backward(ctx, gout):
        img, depth, pose, K, Kinv = ctx.saved_tensors
        rot, pad, align = ctx.cfg
        B, Cc, h, w = img.shape
        gout = gout.contiguous().float()
        gimg = torch.zeros_like(img) if ctx.needs_input_grad[0] else None
        gdepth = torch.empty_like(depth)
        gpose = torch.zeros_like(pose)
        ws = torch.empty(12 * B, dtype=torch.float32, device=img.device)
        L.call('dn_inverse_warp_bwd', L.ptr(img), L.ptr(depth), L.ptr(pose), L.ptr(K), L.ptr(Kinv), B, Cc, h, w, rot, pad,
               align, L.ptr(gout), L.ptr(gimg), L.ptr(gdepth), L.ptr(gpose), L.ptr(ws), L.stream_ptr())
        return gimg, gdepth, gpose, None, None, None, None, None


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros',
                 align_corners=False):
    """Inverse warp a source image to the target image plane (reference inverse_warp.py:160-193).

    img [B,3,H,W], depth [B,H,W], pose [B,6], intrinsics / intrinsics_inv [B,3,3] -> [B,3,H,W]."""
    check_sizes(img, 'img', 'B3HW')
    check_sizes(depth, 'depth', 'BHW')
    check_sizes(pose, 'pose', 'B6')
    check_sizes(intrinsics, 'intrinsics', 'B33')
    check_sizes(intrinsics_inv, 'intrinsics', 'B33')
    assert(intrinsics_inv.size() == intrinsics.size())
    # the reference builds its pixel grid from depth's size (inverse_warp.py:178, set_id_grid): it must equal the image's
    assert tuple(depth.shape[-2:]) == tuple(img.shape[-2:]) and depth.size(0) == img.size(0), 'depth %s vs img %s' % (
        tuple(depth.shape), tuple(img.shape))
    return _InverseWarpFn.apply(img, depth, pose, intrinsics, intrinsics_inv, _ROT[rotation_mode], _PAD[padding_mode],
                                int(bool(align_corners)))
