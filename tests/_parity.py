"""Product (CUDA, through the C ABI) vs oracle (CPU restatement) comparisons on seeded inputs.  Every function
returns a dict of error figures; the pytest files assert on them, tools/gpu_check.py just prints them."""
import math

import torch

import _inputs as I
from _harness import rel, maxabs
from oracle import losses as OL
from oracle import nets as ON

import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import loss_functions as PL
from supervised_dispnet_b200 import inverse_warp as PW

DEV = 'cuda'


def _leaf(t):
    return t.detach().clone().requires_grad_(True)


# ------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------
def l1_case(dataset='kitti', B=4, H=128, W=416, all_invalid=False, seed=0):
    gt = I.sparse_gt(B, H, W, seed, dataset)
    if all_invalid:
        gt[1] = 0
    disp = torch.rand(B, 1, H, W, generator=I._gen(seed + 7)) * 5 + 0.02
    d_o, d_p = _leaf(disp), _leaf(disp.to(DEV))
    lo = OL.l1_loss(gt, [1 / d_o], dataset)
    lp = PL.l1_loss(gt.to(DEV), [1 / d_p], dataset)
    if all_invalid:
        return dict(nan_both=bool(math.isnan(float(lo)) and math.isnan(float(lp))))
    lo.backward()
    lp.backward()
    return dict(loss=abs(float(lp) - float(lo)) / abs(float(lo)), grad=rel(d_p.grad, d_o.grad))


def smooth_case(B=2, H=64, W=96, seed=0):
    maps = [torch.rand(B, 1, H >> s, W >> s, generator=I._gen(seed + s)) * 10 + 0.1 for s in range(4)]
    mo, mp = [_leaf(m) for m in maps], [_leaf(m.to(DEV)) for m in maps]
    lo, lp = OL.smooth_loss(mo), PL.smooth_loss(mp)
    lo.backward()
    lp.backward()
    return dict(loss=abs(float(lp) - float(lo)) / abs(float(lo)), grad=max(rel(a.grad, b.grad) for a, b in zip(mp, mo)))


def smooth_kat():
    H, W = 16, 24
    u = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(1, 1, H, W).contiguous()
    ramp = float(PL.smooth_loss([u.to(DEV)]))
    quad = float(PL.smooth_loss([(u * u).to(DEV)]))
    return dict(ramp=abs(ramp), quad=abs(quad - 2.0))


def _photo_inputs(B, R, H, W, seed, with_mask):
    tgt = I.images(B, H, W, seed)
    refs = [I.images(B, H, W, seed + 1 + i) for i in range(R)]
    K, Kinv = I.intrinsics(B, scale=W / 416.0)
    depth = [I.depth_map(B, H >> s, W >> s, seed + 10 + s).unsqueeze(1) for s in range(4)]
    masks = [I.mask_map(B, R, H >> s, W >> s, seed + 20 + s) for s in range(4)] if with_mask else [None] * 4
    pose = I.poses(B, R, seed + 30)
    return tgt, refs, K, Kinv, depth, masks, pose


def photo_case(rotation_mode='euler', padding_mode='zeros', with_mask=False, B=2, R=2, H=64, W=96, seed=0):
    tgt, refs, K, Kinv, depth, masks, pose = _photo_inputs(B, R, H, W, seed, with_mask)
    d_o, d_p = [_leaf(d) for d in depth], [_leaf(d.to(DEV)) for d in depth]
    m_o = [_leaf(m) if m is not None else None for m in masks]
    m_p = [_leaf(m.to(DEV)) if m is not None else None for m in masks]
    p_o, p_p = _leaf(pose), _leaf(pose.to(DEV))
    lo = OL.photometric_reconstruction_loss(tgt, refs, K, Kinv, d_o, m_o, p_o, rotation_mode, padding_mode)
    lp = PL.photometric_reconstruction_loss(tgt.to(DEV), [r.to(DEV) for r in refs], K.to(DEV), Kinv.to(DEV), d_p, m_p, p_p,
                                            rotation_mode, padding_mode)
    lo.backward()
    lp.backward()
    out = dict(loss=abs(float(lp) - float(lo)) / abs(float(lo)), gdepth=max(rel(a.grad, b.grad) for a, b in zip(d_p, d_o)),
               gpose=rel(p_p.grad, p_o.grad))
    if with_mask:
        out['gmask'] = max(rel(a.grad, b.grad) for a, b in zip(m_p, m_o))
    return out


def warp_case(rotation_mode='euler', padding_mode='zeros', B=2, H=32, W=48, seed=0, identity=False):
    img = I.images(B, H, W, seed)
    K, Kinv = I.intrinsics(B, scale=W / 416.0)
    depth = I.depth_map(B, H, W, seed + 1)
    pose = torch.zeros(B, 6) if identity else I.poses(B, 1, seed + 2)[:, 0]
    i_o, d_o, p_o = _leaf(img), _leaf(depth), _leaf(pose)
    i_p, d_p, p_p = _leaf(img.to(DEV)), _leaf(depth.to(DEV)), _leaf(pose.to(DEV))
    wo = OL.inverse_warp(i_o, d_o, p_o, K, Kinv, rotation_mode, padding_mode)
    wp = PW.inverse_warp(i_p, d_p, p_p, K.to(DEV), Kinv.to(DEV), rotation_mode, padding_mode)
    probe = I.probe_like(wo, seed)
    (wo * probe).sum().backward()
    (wp * probe.to(DEV)).sum().backward()
    # border pixels can flip in/out of view on 1-ulp coordinate differences: report the fraction of differing pixels
    diff = (wp.detach().cpu() - wo.detach()).abs().amax(1)
    return dict(fwd=rel(wp, wo), frac_bad=float((diff > 1e-4).float().mean()), gimg=rel(i_p.grad, i_o.grad),
                gdepth=rel(d_p.grad, d_o.grad), gpose=rel(p_p.grad, p_o.grad))


def errors_case(dataset='kitti', B=4, H=128, W=416, seed=0, unsupervised=False):
    gt = I.sparse_gt(B, H, W, seed, dataset, density=0.05 if dataset == 'kitti' else 0.9)
    pred = torch.rand(B, H, W, generator=I._gen(seed + 3)) * (85 if dataset == 'kitti' else 11)
    crop = True
    eo = OL.compute_errors(gt, pred, dataset, crop, unsupervised)
    ep = PL.compute_errors(gt.to(DEV), pred.to(DEV), dataset, crop, unsupervised)
    out = dict(floats=max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(ep, eo)))
    if not unsupervised:
        co = OL.error_counters(gt, pred, dataset, crop)
        cp, _ = PL.error_counters(gt.to(DEV), pred.to(DEV), dataset, crop)
        out['counters_equal'] = bool((co == cp).all())
        out['n_valid'] = int(co[:, 0].sum())
    return out


def explain_case(B=2, R=2, H=32, W=48, seed=0):
    masks = [I.mask_map(B, R, H >> s, W >> s, seed + s) for s in range(4)]
    mo, mp = [_leaf(m) for m in masks], [_leaf(m.to(DEV)) for m in masks]
    lo, lp = OL.explainability_loss(mo), PL.explainability_loss(mp)
    lo.backward()
    lp.backward()
    return dict(loss=abs(float(lp) - float(lo)) / abs(float(lo)), grad=max(rel(a.grad, b.grad) for a, b in zip(mp, mo)))


def ssim_case(B=2, H=32, W=48, seed=0):
    from supervised_dispnet_b200 import layers as LY
    x = I.images(B, H, W, seed) * 0.5 + 0.5
    y = I.images(B, H, W, seed + 1) * 0.5 + 0.5
    xo, yo, xp, yp = _leaf(x), _leaf(y), _leaf(x.to(DEV)), _leaf(y.to(DEV))
    so, sp = OL.ssim(xo, yo), LY.SSIM()(xp, yp)
    probe = I.probe_like(so, seed)
    (so * probe).sum().backward()
    (sp * probe.to(DEV)).sum().backward()
    return dict(fwd=rel(sp, so), gx=rel(xp.grad, xo.grad), gy=rel(yp.grad, yo.grad))


def edge_smooth_case(B=2, H=32, W=48, seed=0):
    from supervised_dispnet_b200 import layers as LY
    img = I.images(B, H, W, seed) * 0.5 + 0.5
    disp = I.depth_map(B, H, W, seed + 1).unsqueeze(1)
    do, dp = _leaf(disp), _leaf(disp.to(DEV))
    lo, lp = OL.get_smooth_loss(do, img), LY.get_smooth_loss(dp, img.to(DEV))
    lo.backward()
    lp.backward()
    return dict(loss=abs(float(lp) - float(lo)) / abs(float(lo)), grad=rel(dp.grad, do.grad))


def depth_errors_raw_case(seed=0):
    from supervised_dispnet_b200 import layers as LY
    a = I.depth_map(1, 8, 200, seed).flatten()
    b = I.depth_map(1, 8, 200, seed + 1).flatten()
    eo = OL.compute_depth_errors(a, b)
    ep = LY.compute_depth_errors(a.to(DEV), b.to(DEV))
    return dict(floats=max(abs(float(u) - float(v)) / max(abs(float(v)), 1e-12) for u, v in zip(ep, eo)))


LOSS_CASES = [
    ('layers_ssim', ssim_case),
    ('layers_edge_smooth', edge_smooth_case),
    ('layers_depth_errors', depth_errors_raw_case),
    ('l1_kitti', lambda: l1_case('kitti')),
    ('l1_nyu', lambda: l1_case('nyu', H=64, W=80)),
    ('l1_all_invalid_nan', lambda: l1_case('kitti', all_invalid=True)),
    ('smooth_random', smooth_case),
    ('smooth_kat', smooth_kat),
    ('errors_kitti', lambda: errors_case('kitti')),
    ('errors_nyu', lambda: errors_case('nyu', H=64, W=80)),
    ('errors_kitti_median', lambda: errors_case('kitti', unsupervised=True)),
    ('explainability', explain_case),
    ('warp_euler_zeros', lambda: warp_case('euler', 'zeros')),
    ('warp_quat_zeros', lambda: warp_case('quat', 'zeros')),
    ('warp_euler_border', lambda: warp_case('euler', 'border')),
    ('warp_identity', lambda: warp_case('euler', 'zeros', identity=True)),
    ('photo_euler_zeros', lambda: photo_case('euler', 'zeros')),
    ('photo_quat_zeros_mask', lambda: photo_case('quat', 'zeros', with_mask=True)),
    ('photo_euler_border_mask_r4', lambda: photo_case('euler', 'border', with_mask=True, R=4)),
]


# ------------------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------------------
def _load(model, sd):
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all('classifier' in k for k in missing), missing


def _model_compare(prod, oracle_fwd, sd, inputs, precision, n_out):
    prod.precision = precision
    prod.to(DEV).train()
    sd_o = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v.clone())
            for k, v in sd.items()}
    outs_o = oracle_fwd(sd_o, *inputs)
    outs_p = prod(*[x.to(DEV) if torch.is_tensor(x) else [t.to(DEV) for t in x] for x in inputs])
    flat_o, flat_p = _flatten(outs_o), _flatten(outs_p)
    assert len(flat_o) == len(flat_p) == n_out, (len(flat_o), len(flat_p))
    res = dict(out=max(rel(p, o) for p, o in zip(flat_p, flat_o)))
    loss_o = sum((o * I.probe_like(o, i)).sum() for i, o in enumerate(flat_o))
    loss_p = sum((p * I.probe_like(o, i).to(DEV)).sum() for i, (p, o) in enumerate(zip(flat_p, flat_o)))
    loss_o.backward()
    loss_p.backward()
    named = dict(prod.named_parameters())
    gnorm = math.sqrt(sum(float(v.grad.double().norm() ** 2) for k, v in sd_o.items() if v.requires_grad and v.grad is not None))
    errs, num2 = [], 0.0
    for k, v in sd_o.items():
        if not (v.requires_grad and v.grad is not None):
            continue
        g = named[k].grad
        if g is None:
            errs.append((float('inf'), k + ' (missing)'))
            continue
        d = float((g.detach().double().cpu() - v.grad.double()).norm())
        num2 += d * d
        # relative to the parameter's own gradient norm, floored at 1e-3 of the global gradient norm so that
        # analytically-zero gradients (conv biases in front of BatchNorm) are judged on an absolute scale
        errs.append((d / max(float(v.grad.double().norm()), 1e-3 * gnorm), k))
    errs.sort(reverse=True)
    res['grad'] = errs[0][0]
    res['gglobal'] = math.sqrt(num2) / gnorm
    res['worst'] = ','.join('%s:%.1e' % (k, e) for e, k in errs[:3])
    bufs = dict(prod.named_buffers())
    rs = [rel(bufs[k], sd_o[k]) for k in sd_o if 'running' in k]
    if rs:
        res['running'] = max(rs)
    return res


def _flatten(o):
    out = []
    for t in (o if isinstance(o, (list, tuple)) else [o]):
        if isinstance(t, (list, tuple)):
            out += [x for x in t if x is not None]
        elif t is not None:
            out.append(t)
    return out


def vgg_case(precision='fp32', B=2, H=64, W=96):
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    m = S.models.Disp_vgg_BN()
    _load(m, sd)
    return _model_compare(m, lambda s, x: ON.disp_vgg_bn(s, x, True), sd, [I.images(B, H, W, 3)], precision, 4)


def dispnets_case(precision='fp32', B=2, H=128, W=160):
    sd = ON.init_state_dict('DispNetS', 0)
    m = S.models.DispNetS()
    _load(m, sd)
    return _model_compare(m, lambda s, x: ON.dispnets(s, x, True), sd, [I.images(B, H, W, 2)], precision, 4)


def res50_case(precision='fp32', B=2, H=64, W=96):
    sd = ON.init_state_dict('Disp_res_50', 0)
    m = S.models.Disp_res_50()
    _load(m, sd)
    return _model_compare(m, lambda s, x: ON.disp_res_50(s, x, True), sd, [I.images(B, H, W, 4)], precision, 4)


def pose_case(precision='fp32', R=2, exp=False, B=2, H=128, W=160):
    sd = ON.init_state_dict('PoseExpNet', 0, nb_ref_imgs=R, output_exp=exp)
    m = S.models.PoseExpNet(R, exp)
    _load(m, sd)
    tgt = I.images(B, H, W, 5)
    refs = [I.images(B, H, W, 6 + i) for i in range(R)]
    return _model_compare(m, lambda s, t, r: ON.poseexpnet(s, t, r, True, exp), sd, [tgt, refs], precision, 5 if exp else 1)


def vgg_eval_case(precision='fp32', B=1, H=64, W=96):
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    m = S.models.Disp_vgg_BN()
    _load(m, sd)
    m.precision = precision
    m.to(DEV).eval()
    x = I.images(B, H, W, 9)
    with torch.no_grad():
        o = ON.disp_vgg_bn({k: v.clone() for k, v in sd.items()}, x, False)
        p = m(x.to(DEV))
    return dict(out=rel(p, o))


MODEL_CASES = [
    ('Disp_vgg_BN_train', vgg_case),
    ('Disp_vgg_BN_eval', vgg_eval_case),
    ('DispNetS_train', dispnets_case),
    ('PoseExpNet_r2', lambda p: pose_case(p, 2, False)),
    ('PoseExpNet_r4_exp', lambda p: pose_case(p, 4, True)),
    ('Disp_res_50_train', res50_case),
]


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[1] step at full image size: disparities, loss scalar, per-parameter gradients
# ------------------------------------------------------------------------------------------------------------
def _grad_errors(got, want, floor_frac=1e-3):
    """Per-parameter and global L2-relative gradient errors of `got` against `want` (dicts name -> tensor).  A parameter's
    error is taken relative to max(its own gradient norm, floor_frac x the global norm): analytically-zero gradients (conv
    biases in front of BatchNorm) are judged on an absolute scale."""
    gnorm = math.sqrt(sum(float(v.double().norm() ** 2) for v in want.values()))
    num2, per = 0.0, {}
    for k, w in want.items():
        d = float((got[k].detach().double().cpu() - w.double()).norm())
        num2 += d * d
        per[k] = d / max(float(w.double().norm()), floor_frac * gnorm)
    worst = max(per, key=per.get)
    return dict(glob=math.sqrt(num2) / gnorm, worst=per[worst], worst_name=worst, per=per)


def config2_step_case(precision, B=4, H=128, W=416, fp64=False, seed=200):
    """Disp_vgg_BN (train mode) + l1_loss + 0*smooth_loss on B x 3 x H x W (train.py:420-522), product vs oracle (fp32 CPU),
    optionally also vs the oracle evaluated in fp64 -- the yardstick for gradients: the fp32 reference itself is only
    ~1e-3 away from the exact gradient of this 23-layer BatchNorm/ReLU stack."""
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    m = S.models.Disp_vgg_BN()
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    m.precision = precision
    m.to(DEV).train()
    x = I.images(B, H, W, seed=seed)
    gt = I.sparse_gt(B, H, W, seed=seed + 1, dataset='kitti')

    def oracle(dtype):
        s = {k: (v.clone().to(dtype).requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k
                 else (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone())) for k, v in sd.items()}
        d = ON.disp_vgg_bn(s, x.to(dtype), True)
        l = OL.l1_loss(gt.to(dtype), [1 / t for t in d], 'kitti')
        l.backward()
        return [t.detach() for t in d], float(l), {k: v.grad for k, v in s.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}

    d32, l32, g32 = oracle(torch.float32)
    dp = m(x.to(DEV))
    lp = PL.l1_loss(gt.to(DEV), [1 / d for d in dp], 'kitti') + 0.0 * PL.smooth_loss([1 / d for d in dp])
    lp.backward()
    gp = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    e32 = _grad_errors(gp, g32)
    res = dict(disp=[rel(a, b) for a, b in zip(dp, d32)], loss=abs(float(lp) - l32) / abs(l32), grad_global=e32['glob'],
               grad_worst=e32['worst'], grad_worst_name=e32['worst_name'], loss_value=float(lp))
    if fp64:
        d64, l64, g64 = oracle(torch.float64)
        ep, er = _grad_errors(gp, g64), _grad_errors(g32, g64)
        res.update(disp_vs64=[rel(a, b) for a, b in zip(dp, d64)], ref32_disp_vs64=[rel(a, b) for a, b in zip(d32, d64)],
                   grad_global_vs64=ep['glob'], grad_worst_vs64=ep['worst'], grad_worst_name_vs64=ep['worst_name'],
                   ref32_grad_global_vs64=er['glob'], ref32_grad_worst_vs64=er['worst'],
                   worst_ratio=max(ep['per'][k] / max(er['per'][k], 1e-4) for k in g64))
    return res
