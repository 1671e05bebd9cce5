"""Drop-in `loss_functions` module (reference: loss_functions.py): same names, argument meaning and return
types for the functions on the training hot path, each executed by fused CUDA kernels of libdispnet_b200.so.

  l1_loss / l2_loss / berhu_loss / Scale_invariant_loss / Multiscale_{L1,FULL_L1,L2,berhu,scale_inv}_loss
                                    reference :77-315    one masked-reduce kernel family, deterministic, no host sync
  photometric_reconstruction_loss   reference :317-354   area pyramid + one fused warp/photometric kernel per
                                                          (scale, ref) pair, fused analytic backward
  explainability_loss               reference :357-364
  smooth_loss                       reference :367-386   one stencil+reduce kernel per scale
  compute_errors                    reference :401-448   one kernel; integer counters bit-exact

Losses are 0-dim CUDA tensors supporting .item(), arithmetic and .backward() exactly as train.py:488-521 uses
them.  There is no CPU implementation: CPU tensors raise.
"""
import numpy as np
import torch

from . import _lib as L
from .inverse_warp import _PAD, _ROT

# the reference asserts `loss == loss` after every (scale, ref) pair (:342), a host sync each; here the NaN flag
# is accumulated on the device and only read back when this is switched on.
STRICT_NAN_CHECK = False
ALIGN_CORNERS = False     # see inverse_warp.py docstring


def _max_depth(datasets):
    if datasets == 'kitti':
        return 80.0
    if datasets == 'nyu':
        return 10.0
    raise TypeError('undefined datasets')


# ---------------------------------------------------------------------------------------------------------
_KIND = {'l1': 0, 'l2': 1, 'berhu': 2, 'scale_inv': 3}


class _DepthLossFn(torch.autograd.Function):
    """One masked-reduce kernel family for every supervised depth loss (dn_depth_loss_fwd / _bwd): deterministic
    two-stage reductions, no host sync, NaN on an empty mask like `mean()` of an empty selection."""

    @staticmethod
    def forward(ctx, gt, pred, cfg):
        kind, joint, weight, maxd, upf, upmode = cfg
        L.require_cuda(gt, pred)
        gt, pred = gt.contiguous().float(), pred.contiguous().float()
        B, H, W = gt.shape
        assert tuple(pred.shape) == (B, H // upf, W // upf) and H % upf == 0 and W % upf == 0, (tuple(gt.shape), tuple(pred.shape), upf)
        ws = torch.empty(int(L.lib().dn_depth_loss_ws_floats(B)), dtype=torch.float32, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        L.call('dn_depth_loss_fwd', L.ptr(gt), L.ptr(pred), B, H, W, upf, upmode, maxd, kind, int(joint), weight, 0, L.ptr(ws),
               L.ptr(loss), L.stream_ptr())
        ctx.save_for_backward(gt, pred, ws)
        ctx.cfg = cfg
        return loss

    @staticmethod
    def backward(ctx, gout):
        gt, pred, ws = ctx.saved_tensors
        kind, joint, weight, maxd, upf, upmode = ctx.cfg
        B, H, W = gt.shape
        g = torch.empty_like(pred)
        gout = gout.contiguous().float()
        L.call('dn_depth_loss_bwd', L.ptr(gt), L.ptr(pred), B, H, W, upf, upmode, maxd, kind, int(joint), weight, L.ptr(ws),
               L.ptr(gout), L.ptr(g), L.stream_ptr())
        return None, g, None


def _per_sample(kind, gt_depth, depth, datasets):
    maxd = _max_depth(datasets)
    pred = depth[0][:, 0]
    # the reference indexes pred[valid] with a mask built from gt (:112-121): mismatching shapes raise there, and must not
    # become an out-of-bounds device read here
    assert tuple(gt_depth.shape) == tuple(pred.shape), 'gt_depth %s vs depth[0][:, 0] %s' % (tuple(gt_depth.shape), tuple(pred.shape))
    return _DepthLossFn.apply(gt_depth, pred, (_KIND[kind], False, 1.0, maxd, 1, 0))


def l1_loss(gt_depth, depth, datasets):
    """sum_b mean_{valid_b} |gt - clamp(pred, 1e-3, max)| / B, using depth[0][:, 0] only (reference :104-129)."""
    return _per_sample('l1', gt_depth, depth, datasets)


def l2_loss(gt_depth, depth, datasets):
    """reference :77-102.  NOTE the reference's 'nyu' branch computes the mean ABSOLUTE error (:97), not the squared one;
    that behaviour is kept."""
    return _per_sample('l2' if datasets == 'kitti' else 'l1', gt_depth, depth, datasets)


def berhu_loss(gt_depth, depth, datasets):
    """reverse Huber with c = 0.2 max|residual| per sample (reference :131-160).  The reference's 'nyu' branch has lost its
    loop header (:147-159 read `current_gt` before assignment); the same UnboundLocalError is raised here."""
    if datasets == 'nyu':
        raise UnboundLocalError("cannot access local variable 'current_gt' where it is not associated with a value "
                                "(reference loss_functions.py:148)")
    if datasets != 'kitti':          # the reference falls through both branches and returns 0 / B
        return torch.zeros((), dtype=torch.float32, device=depth[0].device)
    return _per_sample('berhu', gt_depth, depth, datasets)


def Scale_invariant_loss(gt_depth, depth, datasets):
    """mean((|gt| - |pred|)^2) - 0.5 (sum(gt - pred))^2 / n^2 per sample (reference :162-187)."""
    if datasets not in ('kitti', 'nyu'):
        return torch.zeros((), dtype=torch.float32, device=depth[0].device)
    return _per_sample('scale_inv', gt_depth, depth, datasets)


def _pool2(x, mode):
    x = x.contiguous().float()
    B, H, W = x.shape
    out = torch.empty((B, H // 2, W // 2), dtype=torch.float32, device=x.device)
    L.call('dn_pool2', L.ptr(x), B, H, W, mode, L.ptr(out), L.stream_ptr())
    return out


def generate_max_pyramid(image):
    """reference :189-194"""
    L.require_cuda(image)
    pyr = [image]
    for i in range(3):
        pyr.append(_pool2(pyr[i], 1))
    return pyr


def generate_avg_pyramid(image):
    """reference :196-201"""
    L.require_cuda(image)
    pyr = [image]
    for i in range(3):
        pyr.append(_pool2(pyr[i], 0))
    return pyr


def generate_bilinear_pyramid(image):
    """reference :203-215: F.interpolate(scale_factor=0.5, mode='bilinear', align_corners=False) of an even-sized map samples
    exactly between the four pixels of every 2x2 block, i.e. it is their mean."""
    return generate_avg_pyramid(image)


def _multiscale(kind, gt_depth, depth, pool_type='bilinear'):
    if pool_type == 'max':
        gts = generate_max_pyramid(gt_depth)
    elif pool_type == 'avg':
        gts = generate_avg_pyramid(gt_depth)
    elif pool_type == 'bilinear':
        gts = generate_bilinear_pyramid(gt_depth)
    else:
        raise TypeError('undefined pool type')
    loss = 0
    for i in range(len(depth)):
        pred = depth[i][:, 0] if depth[i].dim() == 4 else depth[i]
        loss = loss + _DepthLossFn.apply(gts[i], pred, (_KIND[kind], True, 1.0 / (2 ** i), 80.0, 1, 0))
    return loss


def Multiscale_L1_loss(gt_depth, depth, pool_type='bilinear'):
    """sum_i mean_{valid_i} |gt_i - clamp(pred_i)| / 2^i with ONE mask over the whole batch per scale (reference :217-222...)."""
    return _multiscale('l1', gt_depth, depth, pool_type)


def Multiscale_L2_loss(gt_depth, depth):
    """reference :243-258"""
    return _multiscale('l2', gt_depth, depth)


def Multiscale_berhu_loss(gt_depth, depth):
    """reference :260-283"""
    return _multiscale('berhu', gt_depth, depth)


def Multiscale_scale_inv_loss(gt_depth, depth):
    """reference :285-315"""
    return _multiscale('scale_inv', gt_depth, depth)


def Multiscale_FULL_L1_loss(gt_depth, depth, pool_type='bilinear'):
    """every scale's prediction up-sampled x2^i (F.upsample, mode = pool_type) against the full-resolution ground truth
    (reference :224-241); the up-sampling is fused into the masked-reduce kernel and its backward."""
    if pool_type not in ('bilinear', 'nearest'):
        raise NotImplementedError("Multiscale_FULL_L1_loss: up-sampling mode %r" % pool_type)
    loss = 0
    for i in range(len(depth)):
        pred = depth[i][:, 0] if depth[i].dim() == 4 else depth[i]
        loss = loss + _DepthLossFn.apply(gt_depth, pred, (_KIND['l1'], True, 1.0 / (2 ** i), 80.0, 2 ** i, int(pool_type == 'bilinear')))
    return loss


# ---------------------------------------------------------------------------------------------------------
class _SmoothFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *maps):
        L.require_cuda(*maps)
        maps = [m.contiguous().float() for m in maps]
        loss = torch.zeros((), dtype=torch.float32, device=maps[0].device)
        st = L.stream_ptr()
        weight = 1.0
        ctx.weights = []
        for m in maps:
            b, _, h, w = m.shape
            L.call('dn_smooth_fwd', L.ptr(m), b * m.shape[1], h, w, weight, L.ptr(loss), st)
            ctx.weights.append(weight)
            weight /= 2.3
        ctx.save_for_backward(*maps)
        return loss

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous().float()
        st = L.stream_ptr()
        grads = []
        for m, wgt in zip(ctx.saved_tensors, ctx.weights):
            g = torch.empty_like(m)
            b, _, h, w = m.shape
            L.call('dn_smooth_bwd', L.ptr(m), b * m.shape[1], h, w, wgt, L.ptr(gout), L.ptr(g), st)
            grads.append(g)
        return tuple(grads)


def smooth_loss(pred_map):
    """Second-order smoothness over the scale list, weights 1, 1/2.3, ... (reference :367-386)."""
    if type(pred_map) not in [tuple, list]:
        pred_map = [pred_map]
    return _SmoothFn.apply(*pred_map)


# ---------------------------------------------------------------------------------------------------------
class _ExplainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *masks):
        L.require_cuda(*masks)
        masks = [m.contiguous().float() for m in masks]
        loss = torch.zeros((), dtype=torch.float32, device=masks[0].device)
        st = L.stream_ptr()
        for m in masks:
            L.call('dn_explain_fwd', L.ptr(m), m.numel(), L.ptr(loss), st)
        ctx.save_for_backward(*masks)
        return loss

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous().float()
        st = L.stream_ptr()
        grads = []
        for m in ctx.saved_tensors:
            g = torch.empty_like(m)
            L.call('dn_explain_bwd', L.ptr(m), m.numel(), L.ptr(gout), L.ptr(g), st)
            grads.append(g)
        return tuple(grads)


def explainability_loss(mask):
    """sum over scales of BCE(mask, 1) (reference :357-364)."""
    if type(mask) not in [tuple, list]:
        mask = [mask]
    return _ExplainFn.apply(*mask)


# ---------------------------------------------------------------------------------------------------------
def _area_down(x, h, w):
    B, Cc, H, W = x.shape
    if (H, W) == (h, w):
        return x
    f = H // h
    if f * h != H or W // w != f or f * w != W:
        raise RuntimeError('photometric pyramid needs integer area ratios (got %dx%d -> %dx%d)' % (H, W, h, w))
    out = torch.empty((B, Cc, h, w), dtype=torch.float32, device=x.device)
    L.call('dn_area_down', L.ptr(x), B * Cc, H, W, f, L.ptr(out), L.stream_ptr())
    return out


BATCHED_PHOTO = True      # False: one launch per (scale, reference frame) pair (the round-1 path, kept for odd pyramid shapes)


def _photo_batch_ok(H, W, depths, R):
    """The three-launch path serves the reference's own pyramid: every depth map is the full size divided by 1, 2, 4 or 8."""
    if not BATCHED_PHOTO or R > L.PHOTO_MAX_REFS or len(depths) > L.PHOTO_MAX_SCALES or (H % 8) or (W % 8):
        return False
    for d in depths:
        f = H // max(int(d.shape[2]), 1)
        if f not in (1, 2, 4, 8) or d.shape[2] * f != H or d.shape[3] * f != W:
            return False
    return True


class _PhotoBatchFn(torch.autograd.Function):
    """photometric_reconstruction_loss in three launches (dn_area_pyramid, dn_photo_batch_fwd; dn_photo_batch_bwd)."""

    @staticmethod
    def forward(ctx, cfg, tgt, refs, K, Kinv, pose, *maps):
        rot, pad, align, S, has_mask = cfg
        depths = [d.contiguous().float() for d in maps[:S]]
        masks = [m.contiguous().float() for m in maps[S:]] if has_mask else [None] * S
        tgt = tgt.contiguous().float()
        refs = [r.contiguous().float() for r in refs]
        pose = pose.contiguous().float()
        K, Kinv = K.contiguous().float(), Kinv.contiguous().float()
        B, _, H, W = tgt.shape
        R = len(refs)
        dev = tgt.device
        st = L.stream_ptr()
        # /2, /4, /8 area pyramids of the target and every reference frame: one pass over the full-size images
        need = sorted({H // d.shape[2] for d in depths} - {1})
        imgs = [tgt] + refs
        pyr = [{1: im} for im in imgs]
        if need:
            jobs = (L.DnPyrJob * len(imgs))()
            for j, im in enumerate(imgs):
                jobs[j].src = im.data_ptr()
                for f, fld in ((2, 'l1'), (4, 'l2'), (8, 'l3')):
                    if f in need:
                        pyr[j][f] = torch.empty((B, 3, H // f, W // f), dtype=torch.float32, device=dev)
                        setattr(jobs[j], fld, pyr[j][f].data_ptr())
            L.call('dn_area_pyramid', jobs, len(imgs), B * 3, H, W, st)
        P = L.DnPhotoBatch()
        P.nscales, P.nrefs, P.B = S, R, B
        P.rot_mode, P.pad_mode, P.align_corners = rot, pad, align
        P.K, P.Kinv, P.pose = K.data_ptr(), Kinv.data_ptr(), pose.data_ptr()
        for s_, (d, m) in enumerate(zip(depths, masks)):
            f = H // d.shape[2]
            sc = P.sc[s_]
            sc.tgt = pyr[0][f].data_ptr()
            for r in range(R):
                sc.ref[r] = pyr[1 + r][f].data_ptr()
            sc.depth = d.data_ptr()
            sc.mask = m.data_ptr() if m is not None else None
            sc.h, sc.w, sc.downscale = d.shape[2], d.shape[3], float(H) / d.shape[2]
        ws = torch.empty(int(L.lib().dn_photo_ws_floats(C_byref(P))), dtype=torch.float32, device=dev)
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        nanflag = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call('dn_photo_batch_fwd', C_byref(P), L.ptr(ws), L.ptr(loss), L.ptr(nanflag), st)
        if STRICT_NAN_CHECK:
            assert int(nanflag.item()) == 0, 'photometric loss is NaN'
        ctx.cfg = cfg
        ctx.P = P
        ctx.keep = (pyr, depths, masks, pose, K, Kinv, ws)
        ctx.nanflag = nanflag
        return loss

    @staticmethod
    def backward(ctx, gout):
        rot, pad, align, S, has_mask = ctx.cfg
        pyr, depths, masks, pose, K, Kinv, ws = ctx.keep
        P = ctx.P
        gout = gout.contiguous().float()
        gpose = torch.empty_like(pose)
        gdepths = [torch.empty_like(d) for d in depths]
        gmasks = [torch.empty_like(m) if m is not None else None for m in masks]
        for s_ in range(S):
            P.sc[s_].gdepth = gdepths[s_].data_ptr()
            P.sc[s_].gmask = gmasks[s_].data_ptr() if gmasks[s_] is not None else None
        L.call('dn_photo_batch_bwd', C_byref(P), L.ptr(gout), L.ptr(ws), L.ptr(gpose), L.stream_ptr())
        res = [None, None, None, None, None, gpose] + gdepths
        if has_mask:
            res += gmasks
        return tuple(res)


def C_byref(x):
    return L.C.byref(x)


class _PhotoFn(torch.autograd.Function):
    """inputs: pose [B,R,6], then S depth maps [B,1,h,w], then S masks [B,R,h,w] (or absent)."""

    @staticmethod
    def forward(ctx, cfg, tgt, refs, K, Kinv, pose, *maps):
        rot, pad, align, S, has_mask = cfg
        depths = [d.contiguous().float() for d in maps[:S]]
        masks = [m.contiguous().float() for m in maps[S:]] if has_mask else [None] * S
        L.require_cuda(tgt, pose, K, Kinv, *depths)
        tgt = tgt.contiguous().float()
        refs = [r.contiguous().float() for r in refs]
        pose = pose.contiguous().float()
        K, Kinv = K.contiguous().float(), Kinv.contiguous().float()
        B, C3, H, W = tgt.shape
        R = len(refs)
        assert C3 == 3 and all(tuple(r.shape) == tuple(tgt.shape) for r in refs), 'tgt / ref images must be [B,3,H,W]'
        assert pose.dim() == 3 and pose.size(0) == B and pose.size(1) == R and pose.size(2) == 6      # reference :319-320
        assert tuple(K.shape) == (B, 3, 3) and tuple(Kinv.shape) == (B, 3, 3)
        for d, m in zip(depths, masks):
            assert d.dim() == 4 and d.size(0) == B and d.size(1) == 1, 'depth maps must be [B,1,h,w]'
            assert m is None or tuple(m.shape) == (B, R, d.size(2), d.size(3)), 'explainability mask must be [B,R,h,w]'
        dev = tgt.device
        st = L.stream_ptr()
        loss = torch.zeros((), dtype=torch.float32, device=dev)
        nanflag = torch.zeros(1, dtype=torch.int32, device=dev)
        saved = []
        for d, m in zip(depths, masks):
            b, _, h, w = d.shape
            downscale = H / h
            tgt_s = _area_down(tgt, h, w)
            refs_s = [_area_down(r, h, w) for r in refs]
            Ks = torch.cat((K[:, 0:2] / downscale, K[:, 2:]), dim=1).contiguous()
            Kis = torch.cat((Kinv[:, :, 0:2] * downscale, Kinv[:, :, 2:]), dim=2).contiguous()
            for i, r in enumerate(refs_s):
                mp = None if m is None else L.C.c_void_p(m.data_ptr() + 4 * i * h * w)
                L.call('dn_warp_photo_fwd', L.ptr(tgt_s), L.ptr(r), L.ptr(d), L.C.c_void_p(pose.data_ptr() + 4 * 6 * i),
                       6 * R, L.ptr(Ks), L.ptr(Kis), mp, R * h * w, B, h, w, rot, pad, align, None, L.ptr(loss),
                       L.ptr(nanflag), st)
            saved.append((tgt_s, refs_s, Ks, Kis))
        if STRICT_NAN_CHECK:
            assert int(nanflag.item()) == 0, 'photometric loss is NaN'
        ctx.cfg = cfg
        ctx.saved = (saved, depths, masks, pose)
        ctx.nanflag = nanflag
        return loss

    @staticmethod
    def backward(ctx, gout):
        rot, pad, align, S, has_mask = ctx.cfg
        saved, depths, masks, pose = ctx.saved
        gout = gout.contiguous().float()
        B, R = pose.shape[0], pose.shape[1]
        dev = pose.device
        st = L.stream_ptr()
        gpose = torch.zeros_like(pose)
        ws = torch.empty(12 * B, dtype=torch.float32, device=dev)
        gdepths, gmasks = [], []
        for (tgt_s, refs_s, Ks, Kis), d, m in zip(saved, depths, masks):
            b, _, h, w = d.shape
            gd = torch.zeros_like(d)
            gm = torch.empty_like(m) if m is not None else None
            for i, r in enumerate(refs_s):
                mp = None if m is None else L.C.c_void_p(m.data_ptr() + 4 * i * h * w)
                gmp = None if m is None else L.C.c_void_p(gm.data_ptr() + 4 * i * h * w)
                L.call('dn_warp_photo_bwd', L.ptr(tgt_s), L.ptr(r), L.ptr(d), L.C.c_void_p(pose.data_ptr() + 4 * 6 * i),
                       6 * R, L.ptr(Ks), L.ptr(Kis), mp, R * h * w, B, h, w, rot, pad, align, L.ptr(gout), L.ptr(gd),
                       L.C.c_void_p(gpose.data_ptr() + 4 * 6 * i), gmp, R * h * w, L.ptr(ws), st)
            gdepths.append(gd)
            gmasks.append(gm)
        res = [None, None, None, None, None, gpose] + gdepths
        if has_mask:
            res += gmasks
        return tuple(res)


def photometric_reconstruction_loss(tgt_img, ref_imgs, intrinsics, intrinsics_inv, depth, explainability_mask, pose,
                                    rotation_mode='euler', padding_mode='zeros'):
    """sum over scales and reference frames of mean |(tgt_s - warp(ref_s)) * in_view * mask| (reference :317-354)."""
    if type(explainability_mask) not in [tuple, list]:
        explainability_mask = [explainability_mask]
    if type(depth) not in [list, tuple]:
        depth = [depth]
    assert (explainability_mask[0] is None) or (len(explainability_mask) == len(depth))
    has_mask = explainability_mask[0] is not None
    S = len(depth)
    if not has_mask:
        S = min(S, len(explainability_mask))       # zip() semantics of the reference loop (:352)
    cfg = (_ROT[rotation_mode], _PAD[padding_mode], int(ALIGN_CORNERS), S, has_mask)
    maps = list(depth[:S]) + (list(explainability_mask[:S]) if has_mask else [])
    L.require_cuda(tgt_img, pose, intrinsics, intrinsics_inv, *depth[:S])
    B, C3, H, W = tgt_img.shape
    R = len(ref_imgs)
    assert C3 == 3 and all(tuple(r.shape) == tuple(tgt_img.shape) for r in ref_imgs), 'tgt / ref images must be [B,3,H,W]'
    assert pose.dim() == 3 and pose.size(0) == B and pose.size(1) == R and pose.size(2) == 6      # reference :319-320
    assert tuple(intrinsics.shape) == (B, 3, 3) and tuple(intrinsics_inv.shape) == (B, 3, 3)
    for d, m in zip(depth[:S], explainability_mask[:S] if has_mask else [None] * S):
        assert d.dim() == 4 and d.size(0) == B and d.size(1) == 1, 'depth maps must be [B,1,h,w]'
        assert m is None or tuple(m.shape) == (B, R, d.size(2), d.size(3)), 'explainability mask must be [B,R,h,w]'
    fn = _PhotoBatchFn if _photo_batch_ok(H, W, depth[:S], R) else _PhotoFn
    return fn.apply(cfg, tgt_img, list(ref_imgs), intrinsics, intrinsics_inv, pose, *maps)


# ---------------------------------------------------------------------------------------------------------
def error_counters(gt, pred, dataset='kitti', crop=True, unsupervised=False):
    """Raw per-sample results of the metric kernel: (counters int64 [B,4], sums float64 [B,5]) on the host.
    counters = n_valid, n(thresh<1.25), n(<1.25^2), n(<1.25^3); these are the bit-exact integers of SURVEY 8(a9)."""
    L.require_cuda(gt, pred)
    assert gt.dim() == 3 and tuple(gt.shape) == tuple(pred.shape), 'gt %s vs pred %s' % (tuple(gt.shape), tuple(pred.shape))
    gt, pred = gt.contiguous().float(), pred.contiguous().float()
    B, H, W = gt.shape
    if dataset == 'kitti':
        if not crop:
            raise UnboundLocalError("max_depth / crop_mask are unbound in the reference for dataset='kitti', crop=False "
                                    "(loss_functions.py:411-420)")
        maxd = 80.0
        y1, y2 = int(0.40810811 * H), int(0.99189189 * H)
        x1, x2 = int(0.03594771 * W), int(0.96405229 * W)
        use_crop = 1
    else:
        maxd = 10.0
        y1, y2, x1, x2 = 0, H, 0, W
        use_crop = 0
    scale = None
    if unsupervised:
        sc = []
        for g, p in zip(gt, pred):
            valid = (g > 0) & (g < maxd)
            if use_crop:
                cm = torch.zeros_like(valid)
                cm[y1:y2, x1:x2] = True
                valid = valid & cm
            sc.append(torch.median(g[valid]) / torch.median(p[valid].clamp(1e-3, maxd)))
        scale = torch.stack(sc).float().contiguous()
    counters = torch.zeros((B, 4), dtype=torch.int32, device=gt.device)
    sums = torch.zeros((B, 5), dtype=torch.float64, device=gt.device)
    L.call('dn_depth_errors', L.ptr(gt), L.ptr(pred), B, H, W, maxd, use_crop, y1, y2, x1, x2, L.ptr(scale), L.ptr(counters),
           L.ptr(sums), L.stream_ptr())
    return counters.cpu().numpy().astype(np.int64), sums.cpu().numpy()


@torch.no_grad()
def compute_errors(gt, pred, dataset='kitti', crop=True, unsupervised=False):
    """[abs_diff, abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3], each summed over samples / B (reference :401-448)."""
    cnt, s = error_counters(gt, pred, dataset, crop, unsupervised)
    B = gt.size(0)
    n = cnt[:, 0].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        per = np.stack([s[:, 0] / n, s[:, 1] / n, s[:, 2] / n, np.sqrt(s[:, 3] / n), np.sqrt(s[:, 4] / n),
                        cnt[:, 1] / n, cnt[:, 2] / n, cnt[:, 3] / n], 1)
    return [float(v) / B for v in per.sum(0)]
