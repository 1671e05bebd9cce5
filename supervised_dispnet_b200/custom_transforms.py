"""Device-side form of the reference's per-sample input transforms (reference: custom_transforms.py).

train.py:137-142 composes ``RandomHorizontalFlip -> ArrayToTensor -> Normalize`` and runs them in the DataLoader workers on
numpy arrays, one sample at a time; the collated fp32 batch (4 bytes per value) then crosses PCIe.  `DeviceTransform` keeps
the frames as the uint8 HWC arrays the loader reads, sends them as they are (1 byte per value) and applies the same three
steps to the whole batch in one kernel on the GPU (`dn_input_transform`), bit-identical to the reference chain.  The
intrinsics fix-up of the flip (:66, cx -> w - cx) and the mirrored ground truth (:64) are produced alongside.

`Compose`, `Normalize`, `ArrayToTensor`, `RandomHorizontalFlip` keep the reference's names and call signatures for code that
builds the transform list the way train.py does; on this path they only record their parameters -- the arithmetic runs in
`DeviceTransform.__call__`.  `RandomScaleCrop` is commented out of the reference's pipeline (train.py:139) and is not built.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib as L


class Normalize(object):
    def __init__(self, mean, std):
        self.mean, self.std = list(mean), list(std)


class ArrayToTensor(object):
    pass


class RandomHorizontalFlip(object):
    pass


class Compose(object):
    def __init__(self, transforms):
        self.transforms = transforms


class DeviceTransform(object):
    """Batch form of Compose([RandomHorizontalFlip()?, ArrayToTensor(), Normalize(mean, std)]) on the device.

    __call__(frames, gt_depth, intrinsics) with
        frames      list of uint8 tensors / arrays [B,H,W,3] (target first, then the reference frames), host or device
        gt_depth    fp32 [B,H,W] (or [B,k,H,W]) or None
        intrinsics  fp32 [B,3,3] or None
    returns (list of fp32 [B,3,H,W] on the device, gt_depth on the device, intrinsics) with the same per-sample flips applied
    to all three, exactly as the reference applies one transform call per sample."""

    def __init__(self, compose_or_mean=None, std=None, flip=None, device='cuda', rng=None):
        mean = compose_or_mean
        if isinstance(compose_or_mean, Compose):
            mean, std, flip = [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], False
            for t in compose_or_mean.transforms:
                if isinstance(t, Normalize):
                    mean, std = t.mean, t.std
                elif isinstance(t, RandomHorizontalFlip):
                    flip = True
        self.mean = (C.c_float * 4)(*(list(mean) + [0.0] * (4 - len(mean))))
        self.std = (C.c_float * 4)(*(list(std) + [1.0] * (4 - len(std))))
        self.flip = bool(flip)
        self.device = torch.device(device)
        self.rng = rng or random

    def draw_flips(self, B):
        """One random.random() < 0.5 per sample, like RandomHorizontalFlip.__call__ (custom_transforms.py:60)."""
        return [1 if (self.flip and self.rng.random() < 0.5) else 0 for _ in range(B)]

    def __call__(self, frames, gt_depth=None, intrinsics=None, flips=None):
        frames = [torch.as_tensor(np.ascontiguousarray(f)) if not torch.is_tensor(f) else f for f in frames]
        B, H, W, Cc = frames[0].shape
        flips = self.draw_flips(B) if flips is None else list(flips)
        fl = torch.tensor(flips, dtype=torch.int32).to(self.device, non_blocking=True)
        st = L.stream_ptr()
        outs = []
        for f in frames:
            assert f.dtype == torch.uint8 and tuple(f.shape) == (B, H, W, Cc)
            f = f.to(self.device, non_blocking=True).contiguous()
            out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=self.device)
            L.call('dn_input_transform', L.ptr(f), B, H, W, Cc, L.ptr(fl), self.mean, self.std, L.ptr(out), st)
            outs.append(out)
        if gt_depth is not None:
            g = torch.as_tensor(gt_depth).to(self.device, non_blocking=True).contiguous().float()
            assert g.shape[0] == B and g.shape[-1] == W
            go = torch.empty_like(g)
            L.call('dn_flip_rows', L.ptr(g), B, g.numel() // (B * W), W, L.ptr(fl), L.ptr(go), st)
            gt_depth = go
        if intrinsics is not None:
            K = torch.as_tensor(intrinsics).clone().float()
            for b, f in enumerate(flips):
                if f:
                    K[b, 0, 2] = W - K[b, 0, 2]             # custom_transforms.py:66
            intrinsics = K.to(self.device)
        return outs, gt_depth, intrinsics
