"""Drop-in for the three pieces of the reference's `layers.py` (vendored monodepth2) that the north star names:
`SSIM` (:215-245), `get_smooth_loss` (:199-212, edge-aware first-order smoothness) and `compute_depth_errors`
(:248-266).  They are not called by the reference's training loop; they are provided as CUDA kernels with analytic
backward passes so the photometric term can be extended the monodepth2 way (0.85*SSIM + 0.15*L1)."""
import torch
import torch.nn as nn

from . import _lib as L


class _SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        L.require_cuda(x, y)
        x, y = x.contiguous().float(), y.contiguous().float()
        B, Cc, h, w = x.shape
        out = torch.empty_like(x)
        L.call('dn_ssim_fwd', L.ptr(x), L.ptr(y), B * Cc, h, w, L.ptr(out), L.stream_ptr())
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, y = ctx.saved_tensors
        B, Cc, h, w = x.shape
        gout = gout.contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        L.call('dn_ssim_bwd', L.ptr(x), L.ptr(y), L.ptr(gout), B * Cc, h, w, L.ptr(gx), L.ptr(gy), L.stream_ptr())
        return gx, gy


class SSIM(nn.Module):
    """Layer to compute the SSIM loss between a pair of images (reference layers.py:215-245)."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        return _SSIMFn.apply(x, y)


class _EdgeSmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        L.require_cuda(disp, img)
        disp, img = disp.contiguous().float(), img.contiguous().float()
        B, _, h, w = disp.shape
        loss = torch.zeros((), dtype=torch.float32, device=disp.device)
        L.call('dn_edge_smooth_fwd', L.ptr(disp), L.ptr(img), B, img.shape[1], h, w, L.ptr(loss), L.stream_ptr())
        ctx.save_for_backward(disp, img)
        return loss

    @staticmethod
    def backward(ctx, gout):
        disp, img = ctx.saved_tensors
        B, _, h, w = disp.shape
        g = torch.empty_like(disp)
        L.call('dn_edge_smooth_bwd', L.ptr(disp), L.ptr(img), B, img.shape[1], h, w, L.ptr(gout.contiguous().float()), L.ptr(g),
               L.stream_ptr())
        return g, None


def get_smooth_loss(disp, img):
    """Edge-aware smoothness of a disparity image (reference layers.py:199-212); gradient flows to `disp`."""
    assert disp.dim() == 4 and disp.size(1) == 1
    return _EdgeSmoothFn.apply(disp, img)


@torch.no_grad()
def compute_depth_errors(gt, pred):
    """abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 of pre-masked 1-D depth tensors (reference layers.py:248-266)."""
    L.require_cuda(gt, pred)
    gt, pred = gt.contiguous().float().flatten(), pred.contiguous().float().flatten()
    n = gt.numel()
    counters = torch.zeros(3, dtype=torch.int32, device=gt.device)
    sums = torch.zeros(4, dtype=torch.float64, device=gt.device)
    L.call('dn_depth_errors_raw', L.ptr(gt), L.ptr(pred), n, L.ptr(counters), L.ptr(sums), L.stream_ptr())
    c, s = counters.double(), sums
    abs_rel, sq_rel = s[0] / n, s[1] / n
    rmse, rmse_log = torch.sqrt(s[2] / n), torch.sqrt(s[3] / n)
    return tuple(v.float() for v in (abs_rel, sq_rel, rmse, rmse_log, c[0] / n, c[1] / n, c[2] / n))
