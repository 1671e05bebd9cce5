"""Seeded synthetic inputs shared by oracle/make_golden.py, the tests and bench.py.

Everything is generated on CPU with a private torch.Generator so a (shape, seed) pair names one tensor
everywhere; callers move the result to the device they need.
"""
import torch


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def images(b, h, w, seed=0, c=3):
    """fp32 NCHW, U(0,1) normalised with mean .5 / std .5 -> [-1, 1] (mirrors train.py:131-132)."""
    return (torch.rand(b, c, h, w, generator=_gen(seed)) - 0.5) / 0.5


def probe_like(t, seed):
    """Fixed random cotangent used to turn a tensor output into a scalar for gradient parity."""
    return torch.randn(t.shape, generator=_gen(1000 + seed))


def depth_map(b, h, w, seed=0, lo=1.0, hi=60.0):
    """Smooth-ish positive depth [B,H,W]: a tilted plane plus noise, so warps stay mostly in view."""
    g = _gen(seed)
    base = torch.rand(b, 1, 1, generator=g) * (hi - lo) * 0.5 + lo
    noise = torch.rand(b, h, w, generator=g) * (hi - lo) * 0.5
    return (base + noise).contiguous()


def mask_map(b, r, h, w, seed=0):
    return torch.rand(b, r, h, w, generator=_gen(seed)) * 0.98 + 0.01


def poses(b, r, seed=0, scale=0.05):
    """[B,R,6] small 6-DoF poses (tx,ty,tz,rx,ry,rz), the size PoseExpNet emits early in training x5."""
    return torch.randn(b, r, 6, generator=_gen(seed)) * scale


def intrinsics(b, scale=1.0):
    """KITTI-like K at 128x416, scaled by `scale` for smaller test images; returns (K, K^-1) [B,3,3]."""
    K = torch.tensor([[241.67 * scale, 0.0, 204.17 * scale], [0.0, 246.28 * scale, 59.0 * scale], [0.0, 0.0, 1.0]])
    K = K.unsqueeze(0).repeat(b, 1, 1)
    return K, torch.inverse(K)


def sparse_gt(b, h, w, seed=0, dataset='kitti', density=0.05):
    """KITTI-like sparse velodyne depth (zeros where missing) or NYU-like dense depth; includes values
    above max_depth so the `< max` test is exercised, and >=1 valid pixel per sample inside the Garg crop."""
    g = _gen(seed)
    hi = 90.0 if dataset == 'kitti' else 11.0
    d = torch.rand(b, h, w, generator=g) * (hi - 0.5) + 0.5
    keep = torch.rand(b, h, w, generator=g) < density
    d = d * keep
    d[:, int(0.7 * h), w // 2] = 5.0
    return d.contiguous()


def subsample(t, n=8192):
    """Deterministic strided subsample used to keep large gradient fixtures small."""
    f = t.detach().flatten()
    if f.numel() <= 2 * n:
        return f.clone()
    return f[:: f.numel() // n][:n].clone()
