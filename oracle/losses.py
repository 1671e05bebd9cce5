"""Oracle (TEST INFRASTRUCTURE ONLY): fp32 restatement of the per-pixel warp / loss / metric functions.

Written from the formulas of SURVEY.md appendix A, *not* by calling the library composites the reference
calls (no ``F.grid_sample``, no ``F.interpolate``): every step is explicit tensor arithmetic so the CUDA
kernels have an independent statement of the same maths to be checked against.  torch autograd through
these functions is the gradient oracle.

Reference lines restated:
  inverse_warp                      inverse_warp.py:160-193  (pixel2cam :26-40, cam2pixel :43-74,
                                    euler2mat :77-114, quat2mat :117-138, pose_vec2mat :141-157)
  photometric_reconstruction_loss   loss_functions.py:317-354
  explainability_loss               loss_functions.py:357-364
  smooth_loss                       loss_functions.py:367-386
  l1_loss                           loss_functions.py:104-129
  l2 / berhu / Scale_invariant / Multiscale_*       loss_functions.py:77-315
  compute_errors                    loss_functions.py:401-448
  SSIM / get_smooth_loss / compute_depth_errors    layers.py:215-245 / :199-212 / :248-266
"""
import numpy as np
import torch


# --------------------------------------------------------------------------------------------------
# pose -> matrix
# --------------------------------------------------------------------------------------------------

def euler2mat(angle):
    """R = Rx(rx) @ Ry(ry) @ Rz(rz)  (inverse_warp.py:95-113)."""
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    o, i = torch.zeros_like(x), torch.ones_like(x)
    cz, sz, cy, sy, cx, sx = z.cos(), z.sin(), y.cos(), y.sin(), x.cos(), x.sin()
    zm = torch.stack([cz, -sz, o, sz, cz, o, o, o, i], 1).view(-1, 3, 3)
    ym = torch.stack([cy, o, sy, o, i, o, -sy, o, cy], 1).view(-1, 3, 3)
    xm = torch.stack([i, o, o, o, cx, -sx, o, sx, cx], 1).view(-1, 3, 3)
    return xm @ ym @ zm


def quat2mat(quat):
    """q = normalise([1, qx, qy, qz]) -> rotation (inverse_warp.py:125-137)."""
    q = torch.cat([torch.ones_like(quat[:, :1]), quat], 1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1).view(-1, 3, 3)


def pose_vec2mat(vec, rotation_mode='euler'):
    rot = euler2mat(vec[:, 3:]) if rotation_mode == 'euler' else quat2mat(vec[:, 3:])
    return torch.cat([rot, vec[:, :3].unsqueeze(-1)], 2)


# --------------------------------------------------------------------------------------------------
# inverse warp with an explicit bilinear gather
# --------------------------------------------------------------------------------------------------

def warp_coords(depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros'):
    """Normalised sampling grid [B,H,W,2] (steps 1-5 of appendix A.1)."""
    b, h, w = depth.shape
    v, u = torch.meshgrid(torch.arange(h, dtype=depth.dtype), torch.arange(w, dtype=depth.dtype), indexing='ij')
    pix = torch.stack([u, v, torch.ones_like(u)], 0).view(1, 3, -1).to(depth.device)      # (u, v, 1)
    cam = (intrinsics_inv @ pix.expand(b, 3, -1)) * depth.reshape(b, 1, -1)                # pixel2cam :38-40
    proj = intrinsics @ pose_vec2mat(pose, rotation_mode)                                  # :188
    p = proj[:, :, :3] @ cam + proj[:, :, 3:]                                              # cam2pixel :55,60
    X, Y, Z = p[:, 0], p[:, 1], p[:, 2].clamp(min=1e-3)
    xn = 2 * (X / Z) / (w - 1) - 1
    yn = 2 * (Y / Z) / (h - 1) - 1
    if padding_mode == 'zeros':
        # masked in-place fill on a non-leaf: no gradient through the replaced coordinates (:67-71)
        xn = torch.where(((xn > 1) | (xn < -1)).detach(), torch.full_like(xn, 2.0), xn)
        yn = torch.where(((yn > 1) | (yn < -1)).detach(), torch.full_like(yn, 2.0), yn)
    return torch.stack([xn, yn], 2).view(b, h, w, 2)


def bilinear_sample(img, grid, padding_mode='zeros', align_corners=False):
    """Restatement of grid_sample(bilinear) semantics (appendix A.1 step 6)."""
    b, c, h, w = img.shape
    xn, yn = grid[..., 0], grid[..., 1]
    if align_corners:
        ix = (xn + 1) / 2 * (w - 1)
        iy = (yn + 1) / 2 * (h - 1)
    else:
        ix = ((xn + 1) * w - 1) / 2
        iy = ((yn + 1) * h - 1) / 2
    if padding_mode == 'border':
        ix = ix.clamp(0, w - 1)
        iy = iy.clamp(0, h - 1)
    x0 = ix.floor()
    y0 = iy.floor()
    x1, y1 = x0 + 1, y0 + 1
    wx1, wy1 = ix - x0, iy - y0
    wx0, wy0 = x1 - ix, y1 - iy
    flat = img.reshape(b, c, h * w)
    out = 0
    for xc, yc, wgt in ((x0, y0, wx0 * wy0), (x1, y0, wx1 * wy0), (x0, y1, wx0 * wy1), (x1, y1, wx1 * wy1)):
        inb = (xc >= 0) & (xc <= w - 1) & (yc >= 0) & (yc <= h - 1)
        idx = (yc.clamp(0, h - 1) * w + xc.clamp(0, w - 1)).long().view(b, 1, -1).expand(b, c, -1)
        val = flat.gather(2, idx).view(b, c, *xn.shape[1:])
        out = out + val * (wgt * inb.to(img.dtype)).unsqueeze(1)
    return out


def inverse_warp(img, depth, pose, intrinsics, intrinsics_inv, rotation_mode='euler', padding_mode='zeros',
                 align_corners=False):
    grid = warp_coords(depth, pose, intrinsics, intrinsics_inv, rotation_mode, padding_mode)
    return bilinear_sample(img, grid, padding_mode, align_corners)


# --------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------

def area_downsample(img, h, w):
    """F.interpolate(mode='area') for integer ratios = block mean (loss_functions.py:326-327)."""
    b, c, H, W = img.shape
    fy, fx = H // h, W // w
    assert fy * h == H and fx * w == W
    return img.view(b, c, h, fy, w, fx).mean(dim=(3, 5))


def photometric_reconstruction_loss(tgt_img, ref_imgs, intrinsics, intrinsics_inv, depth, explainability_mask,
                                    pose, rotation_mode='euler', padding_mode='zeros', align_corners=False):
    if not isinstance(explainability_mask, (list, tuple)):
        explainability_mask = [explainability_mask]
    if not isinstance(depth, (list, tuple)):
        depth = [depth]
    loss = 0
    for d, mask in zip(depth, explainability_mask):
        b, _, h, w = d.shape
        downscale = tgt_img.size(2) / h
        tgt_s = area_downsample(tgt_img, h, w)
        K = torch.cat((intrinsics[:, 0:2] / downscale, intrinsics[:, 2:]), 1)
        Kinv = torch.cat((intrinsics_inv[:, :, 0:2] * downscale, intrinsics_inv[:, :, 2:]), 2)
        for i, ref in enumerate(ref_imgs):
            warped = inverse_warp(area_downsample(ref, h, w), d[:, 0], pose[:, i], K, Kinv, rotation_mode,
                                  padding_mode, align_corners)
            oob = 1 - (warped == 0).all(dim=1, keepdim=True).to(warped.dtype)
            diff = (tgt_s - warped) * oob
            if mask is not None:
                diff = diff * mask[:, i:i + 1]
            loss = loss + diff.abs().mean()
    return loss


def explainability_loss(mask):
    if not isinstance(mask, (list, tuple)):
        mask = [mask]
    loss = 0
    for m in mask:
        loss = loss + (-torch.log(m).clamp(min=-100)).mean()      # BCE(m, 1), torch clamps log at -100
    return loss


def smooth_loss(pred_map):
    if not isinstance(pred_map, (list, tuple)):
        pred_map = [pred_map]
    loss, weight = 0, 1.0
    for p in pred_map:
        dx = p[:, :, :, 1:] - p[:, :, :, :-1]
        dy = p[:, :, 1:] - p[:, :, :-1]
        dx2 = dx[:, :, :, 1:] - dx[:, :, :, :-1]
        dxdy = dx[:, :, 1:] - dx[:, :, :-1]
        dydx = dy[:, :, :, 1:] - dy[:, :, :, :-1]
        dy2 = dy[:, :, 1:] - dy[:, :, :-1]
        loss = loss + (dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()) * weight
        weight /= 2.3
    return loss


def max_depth_of(datasets):
    if datasets == 'kitti':
        return 80.0
    if datasets == 'nyu':
        return 10.0
    raise TypeError('undefined datasets')


def l1_loss(gt_depth, depth, datasets):
    """Per-sample masked mean |gt - clamp(pred)|, summed over samples / B.  Empty mask -> NaN (0/0)."""
    M = max_depth_of(datasets)
    pred = depth[0][:, 0]
    valid = ((gt_depth > 0) & (gt_depth < M)).to(pred.dtype)
    err = (gt_depth - pred.clamp(1e-3, M)).abs() * valid
    per = err.flatten(1).sum(1) / valid.flatten(1).sum(1)
    return per.sum() / pred.size(0)


# --------------------------------------------------------------------------------------------------
# the other supervised depth losses behind train.py's --loss switch (loss_functions.py:77-315), restated with masked sums
# --------------------------------------------------------------------------------------------------
def _masked(gt, pred, M):
    """valid mask (as float), residual gt - clamp(pred) zeroed outside the mask, count per leading index."""
    valid = ((gt > 0) & (gt < M)).to(pred.dtype)
    d = (gt - pred.clamp(1e-3, M)) * valid
    return valid, d


def _berhu_rows(valid, d):
    """rows = samples (or one row for the whole batch): reverse Huber with c = 0.2 max|d| over the row's valid pixels."""
    r = d.abs().flatten(1)
    v = valid.flatten(1)
    c = 0.2 * r.max(1, keepdim=True)[0]
    per = torch.where(r > c, (r * r + c * c) / (2 * c), r) * v
    return per.sum(1) / v.sum(1)


def l2_loss(gt_depth, depth, datasets):
    """loss_functions.py:77-102 (the 'nyu' branch averages |d|, not d^2: line 97)."""
    M = max_depth_of(datasets)
    pred = depth[0][:, 0]
    valid, d = _masked(gt_depth, pred, M)
    e = d * d if datasets == 'kitti' else d.abs()
    return (e.flatten(1).sum(1) / valid.flatten(1).sum(1)).sum() / pred.size(0)


def berhu_loss(gt_depth, depth, datasets):
    """loss_functions.py:131-160, 'kitti' branch (the 'nyu' branch of the reference raises UnboundLocalError)."""
    pred = depth[0][:, 0]
    valid, d = _masked(gt_depth, pred, max_depth_of(datasets))
    return _berhu_rows(valid, d).sum() / pred.size(0)


def _scale_inv_rows(valid, d):
    n = valid.flatten(1).sum(1)
    return (d * d).flatten(1).sum(1) / n - 0.5 * d.flatten(1).sum(1) ** 2 / (n * n)


def scale_invariant_loss(gt_depth, depth, datasets):
    """loss_functions.py:162-187 (gt > 0 and pred >= 1e-3 inside the mask, so |gt| - |pred| = gt - pred)."""
    pred = depth[0][:, 0]
    valid, d = _masked(gt_depth, pred, max_depth_of(datasets))
    return _scale_inv_rows(valid, d).sum() / pred.size(0)


def gt_pyramid(gt, pool_type='bilinear'):
    """loss_functions.py:189-215: three 2x2/stride-2 poolings; 'bilinear' (scale 0.5, align_corners=False) = 'avg' = block mean."""
    pyr = [gt]
    for _ in range(3):
        g = pyr[-1]
        b, h, w = g.shape
        blk = g[:, :h // 2 * 2, :w // 2 * 2].reshape(b, h // 2, 2, w // 2, 2)
        pyr.append(blk.amax((2, 4)) if pool_type == 'max' else blk.mean((2, 4)))
    return pyr


def multiscale_loss(kind, gt_depth, depth, pool_type='bilinear'):
    """Multiscale_{L1,L2,berhu,scale_inv}_loss (loss_functions.py:217-222, :243-315): one mask over the whole batch per scale,
    max depth 80, weight 1/2^i."""
    gts = gt_pyramid(gt_depth, pool_type)
    loss = 0
    for i, dmap in enumerate(depth):
        pred = dmap[:, 0]
        valid, d = _masked(gts[i], pred, 80.0)
        valid, d = valid.reshape(1, -1), d.reshape(1, -1)
        if kind == 'l1':
            v = d.abs().sum() / valid.sum()
        elif kind == 'l2':
            v = (d * d).sum() / valid.sum()
        elif kind == 'berhu':
            v = _berhu_rows(valid, d)[0]
        else:
            v = _scale_inv_rows(valid, d)[0]
        loss = loss + v / (2 ** i)
    return loss


def upsample_bilinear(x, f):
    """F.upsample(x, scale_factor=f, mode='bilinear') (align_corners=False) of [B,h,w] with explicit gathers."""
    b, h, w = x.shape
    ys = ((torch.arange(h * f, dtype=x.dtype) + 0.5) / f - 0.5).clamp(min=0)
    xs = ((torch.arange(w * f, dtype=x.dtype) + 0.5) / f - 0.5).clamp(min=0)
    y0, x0 = ys.floor().long(), xs.floor().long()
    y1, x1 = (y0 + 1).clamp(max=h - 1), (x0 + 1).clamp(max=w - 1)
    ly, lx = (ys - y0.to(x.dtype)).view(1, -1, 1), (xs - x0.to(x.dtype)).view(1, 1, -1)
    g = lambda yy, xx: x[:, yy][:, :, xx]          # noqa: E731
    return (1 - ly) * ((1 - lx) * g(y0, x0) + lx * g(y0, x1)) + ly * ((1 - lx) * g(y1, x0) + lx * g(y1, x1))


def multiscale_full_l1_loss(gt_depth, depth, pool_type='bilinear'):
    """Multiscale_FULL_L1_loss (loss_functions.py:224-241): predictions up-sampled to full size."""
    loss = 0
    for i, dmap in enumerate(depth):
        f = 2 ** i
        pred = dmap[:, 0]
        if f > 1:
            pred = upsample_bilinear(pred, f) if pool_type == 'bilinear' else pred.repeat_interleave(f, 1).repeat_interleave(f, 2)
        valid, d = _masked(gt_depth, pred, 80.0)
        loss = loss + (d.abs().sum() / valid.sum()) / f
    return loss


def resize_bilinear_ac(x, size):
    """nn.UpsamplingBilinear2d(size) = bilinear with align_corners=True on [B,h,w] (train.py:698-700), explicit gathers."""
    b, h, w = x.shape
    H, W = size
    ys = torch.arange(H, dtype=x.dtype) * ((h - 1) / (H - 1) if H > 1 else 0.0)
    xs = torch.arange(W, dtype=x.dtype) * ((w - 1) / (W - 1) if W > 1 else 0.0)
    y0, x0 = ys.floor().long().clamp(max=h - 1), xs.floor().long().clamp(max=w - 1)
    y1, x1 = (y0 + 1).clamp(max=h - 1), (x0 + 1).clamp(max=w - 1)
    ly, lx = (ys - y0.to(x.dtype)).view(1, -1, 1), (xs - x0.to(x.dtype)).view(1, 1, -1)
    g = lambda yy, xx: x[:, yy][:, :, xx]          # noqa: E731
    return (1 - ly) * ((1 - lx) * g(y0, x0) + lx * g(y0, x1)) + ly * ((1 - lx) * g(y1, x0) + lx * g(y1, x1))


def validate_with_gt(forward_eval, loader, dataset):
    """validate_with_gt (train.py:642-723): per batch depth = 1/disp[:, 0] of the eval-mode forward (NYU: resized to the ground
    truth), compute_errors, then the batch average (AverageMeter with n = 1 per batch, logger.py:62-89)."""
    tot, n = None, 0
    with torch.no_grad():
        for x, depth in loader:
            if dataset == 'nyu':
                depth = torch.squeeze(depth[:, 0])
            out = 1 / forward_eval(x)[:, 0]
            if dataset == 'nyu':
                out = resize_bilinear_ac(out, depth.shape[1:])
            e = compute_errors(depth, out, dataset)
            tot = e if tot is None else [a + b for a, b in zip(tot, e)]
            n += 1
    return [t / n for t in tot]


def garg_crop(h, w):
    return int(0.40810811 * h), int(0.99189189 * h), int(0.03594771 * w), int(0.96405229 * w)


def error_counters(gt, pred, dataset='kitti', crop=True):
    """Integer part of compute_errors (numpy): per-sample [n_valid, n<1.25, n<1.25^2, n<1.25^3] as int64.
    The thresholds are Python doubles compared against fp32 tensors, i.e. compared in fp32 after
    rounding the constant to fp32 (appendix A.5)."""
    g = gt.detach().cpu().numpy().astype(np.float32)
    p = pred.detach().cpu().numpy().astype(np.float32)
    B, H, W = g.shape
    M = np.float32(80.0 if dataset == 'kitti' else 10.0)
    cm = np.ones((H, W), bool)
    if dataset == 'kitti' and crop:
        y1, y2, x1, x2 = garg_crop(H, W)
        cm[:] = False
        cm[y1:y2, x1:x2] = True
    out = np.zeros((B, 4), np.int64)
    for b in range(B):
        valid = (g[b] > 0) & (g[b] < M)
        if crop:
            valid &= cm
        vg = g[b][valid]
        vp = np.clip(p[b][valid], np.float32(1e-3), M)
        th = np.maximum(vg / vp, vp / vg)
        out[b] = [valid.sum(), (th < np.float32(1.25)).sum(), (th < np.float32(1.25 ** 2)).sum(),
                  (th < np.float32(1.25 ** 3)).sum()]
    return out


@torch.no_grad()
def compute_errors(gt, pred, dataset='kitti', crop=True, unsupervised=False):
    B, H, W = gt.shape
    M = max_depth_of('kitti' if dataset == 'kitti' else 'nyu')
    crop_mask = torch.ones(H, W, dtype=torch.bool, device=gt.device)
    if dataset == 'kitti' and crop:
        y1, y2, x1, x2 = garg_crop(H, W)
        crop_mask[:] = False
        crop_mask[y1:y2, x1:x2] = True
    acc = [0.0] * 8
    for g, p in zip(gt, pred):
        valid = (g > 0) & (g < M)
        if crop:
            valid = valid & crop_mask
        vg = g[valid]
        vp = p[valid].clamp(1e-3, M)
        if unsupervised:
            vp = vp * torch.median(vg) / torch.median(vp)
        th = torch.max(vg / vp, vp / vg)
        d = vg - vp
        vals = [d.abs().mean(), (d.abs() / vg).mean(), (d * d / vg).mean(), (d * d).mean().sqrt(),
                ((vg.log() - vp.log()) ** 2).mean().sqrt(),
                (th < 1.25).float().mean(), (th < 1.25 ** 2).float().mean(), (th < 1.25 ** 3).float().mean()]
        acc = [a + v for a, v in zip(acc, vals)]
    return [float(a) / B for a in acc]


# --------------------------------------------------------------------------------------------------
# monodepth2-style optional terms named by the north star (layers.py, uncalled in the reference)
# --------------------------------------------------------------------------------------------------

def _reflect_pad1(x):
    x = torch.cat([x[:, :, 1:2], x, x[:, :, -2:-1]], 2)
    return torch.cat([x[:, :, :, 1:2], x, x[:, :, :, -2:-1]], 3)


def _avg3(x):
    h, w = x.shape[2] - 2, x.shape[3] - 2
    s = 0
    for dy in range(3):
        for dx in range(3):
            s = s + x[:, :, dy:dy + h, dx:dx + w]
    return s / 9


def ssim(x, y):
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x, y = _reflect_pad1(x), _reflect_pad1(y)
    mx, my = _avg3(x), _avg3(y)
    sx = _avg3(x * x) - mx * mx
    sy = _avg3(y * y) - my * my
    sxy = _avg3(x * y) - mx * my
    n = (2 * mx * my + C1) * (2 * sxy + C2)
    d = (mx * mx + my * my + C1) * (sx + sy + C2)
    return ((1 - n / d) / 2).clamp(0, 1)


def get_smooth_loss(disp, img):
    gdx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gdy = (disp[:, :, :-1] - disp[:, :, 1:]).abs()
    gix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    giy = (img[:, :, :-1] - img[:, :, 1:]).abs().mean(1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def compute_depth_errors(gt, pred):
    th = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = [(th < 1.25 ** k).float().mean() for k in (1, 2, 3)]
    d = gt - pred
    return ((d.abs() / gt).mean(), (d * d / gt).mean(), (d * d).mean().sqrt(),
            ((gt.log() - pred.log()) ** 2).mean().sqrt(), a1, a2, a3)
