"""CPU: the oracle restatement against fixtures produced by the reference itself (oracle/make_golden.py).

Tolerances: the oracle restates the same fp32 maths with a different op order (explicit gather instead
of grid_sample, masked sums instead of boolean-index means), so agreement is to fp32 round-off:
1e-5 relative on tensors (L2), 1e-5 on scalars; integer counters bit-exact.
"""
import math

import pytest
import torch

import _inputs as I
from oracle import losses as OL
from oracle import nets as ON


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / b.double().norm().clamp_min(1e-30))


def checksum(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()
            if v.dtype.is_floating_point}


@pytest.mark.parametrize('name,model,kw', [
    ('DispNetS', 'DispNetS', {}), ('Disp_res_50', 'Disp_res_50', {}),
    ('PoseExpNet_r2', 'PoseExpNet', dict(nb_ref_imgs=2, output_exp=False)),
    ('PoseExpNet_r4e', 'PoseExpNet', dict(nb_ref_imgs=4, output_exp=True)),
    ('Disp_vgg_BN', 'Disp_vgg_BN', {})])
def test_g0_init_matches_reference_bitwise(golden, name, model, kw):
    fp = golden('g0_init_fingerprints')[name]
    mine = checksum(ON.init_state_dict(model, 0, **kw))
    assert set(mine) == set(fp)
    for k in fp:
        assert mine[k] == fp[k], k


def test_g1_dispnets_eval_config1(golden):
    """BASELINE config 1: DispNetS forward on 1x3x128x416, CPU, disparity vs reference."""
    sd = ON.init_state_dict('DispNetS', 0)
    with torch.no_grad():
        d = ON.dispnets(sd, I.images(1, 128, 416, seed=1), training=False)
    assert rel(d, golden('g1_dispnets_eval')) < 1e-5


def _grad_check(params, loss, gold, tol=2e-5):
    loss.backward()
    for k, g in gold.items():
        assert rel(I.subsample(params[k].grad), g) < tol, k


def test_g1_dispnets_train(golden):
    g = golden('g1_dispnets_train')
    sd = ON.init_state_dict('DispNetS', 0)
    for v in sd.values():
        v.requires_grad_(True)
    outs = ON.dispnets(sd, I.images(2, 128, 160, seed=2), training=True)
    for o, r in zip(outs, g['outs']):
        assert o.shape == r.shape and rel(o, r) < 1e-5
    _grad_check(sd, sum((o * I.probe_like(o, 10 + i)).sum() for i, o in enumerate(outs)), g['grads'])


def test_g2_vgg_train_and_running_stats(golden):
    g = golden('g2_vgg_train')
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    for v in sd.values():
        if v.dtype.is_floating_point and 'running' not in str(id(v)):
            pass
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    outs = ON.disp_vgg_bn(sd, I.images(2, 64, 96, seed=3), training=True)
    for o, r in zip(outs, g['outs']):
        assert o.shape == r.shape and rel(o, r) < 1e-5
    for k, r in g['running'].items():
        if r.dtype.is_floating_point:
            assert rel(sd[k], r) < 1e-5, k
        else:
            assert int(sd[k]) == int(r), k
    _grad_check(params, sum((o * I.probe_like(o, 20 + i)).sum() for i, o in enumerate(outs)), g['grads'], 5e-5)
    assert all(k.startswith('features.classifier') for k in g['no_grad_keys'])


def test_g2_vgg_eval(golden):
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    with torch.no_grad():
        d = ON.disp_vgg_bn(sd, I.images(1, 128, 416, seed=4), training=False)
    assert rel(d, golden('g2_vgg_eval')) < 1e-5


def test_g2b_res50_train(golden):
    g = golden('g2b_res50_train')
    sd = ON.init_state_dict('Disp_res_50', 0)
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    outs = ON.disp_res_50(sd, I.images(2, 64, 96, seed=5), training=True)
    for o, r in zip(outs, g['outs']):
        assert o.shape == r.shape and rel(o, r) < 2e-5
    for k, r in g['running'].items():
        assert rel(sd[k], r) < 1e-5, k
    _grad_check(params, sum((o * I.probe_like(o, 30 + i)).sum() for i, o in enumerate(outs)), g['grads'], 1e-4)
    assert sorted(g['no_grad_keys']) == ['bn1.bias', 'bn1.weight']


@pytest.mark.parametrize('tag,R,exp', [('r2', 2, False), ('r4e', 4, True)])
def test_g2c_poseexpnet(golden, tag, R, exp):
    g = golden('g2c_pose_' + tag)
    sd = ON.init_state_dict('PoseExpNet', 0, nb_ref_imgs=R, output_exp=exp)
    for v in sd.values():
        v.requires_grad_(True)
    hw = (64, 128) if exp else (128, 416)
    tgt = I.images(2, hw[0], hw[1], seed=6)
    refs = [I.images(2, hw[0], hw[1], seed=7 + r) for r in range(R)]
    masks, pose = ON.poseexpnet(sd, tgt, refs, True, exp)
    assert rel(pose, g['pose']) < 1e-5
    loss = (pose * I.probe_like(pose, 40)).sum()
    for i, (m, r) in enumerate(zip(masks, g['masks'])):
        assert (m is None) == (r is None)
        if m is not None:
            assert rel(m, r) < 1e-5
            loss = loss + (m * I.probe_like(m, 41 + i)).sum()
    _grad_check(sd, loss, g['grads'])


def test_g3_inverse_warp(golden):
    g3 = golden('g3_inverse_warp')
    B, h, w = 2, 32, 104
    img = I.images(B, h, w, seed=50)
    K, Kinv = I.intrinsics(B, h / 128.0)
    for pname, pose in (('identity', torch.zeros(B, 6)), ('random', I.poses(B, 1, seed=51)[:, 0])):
        for rot in ('euler', 'quat'):
            for pad in ('zeros', 'border'):
                g = g3['%s_%s_%s' % (pname, rot, pad)]
                depth = I.depth_map(B, h, w, seed=52).requires_grad_(True)
                p = pose.clone().requires_grad_(True)
                out = OL.inverse_warp(img, depth, p, K, Kinv, rot, pad)
                # ULP-level coordinate differences may flip a handful of border pixels (hard part 5)
                bad = ((out - g['out']).abs() > 1e-4).float().mean().item()
                assert bad < 2e-3, (pname, rot, pad, bad)
                (out * I.probe_like(out, 53)).sum().backward()
                if pname == 'identity':     # t = 0: depth cancels analytically, the gradient is round-off noise
                    assert float(depth.grad.norm()) < 1e-3 and float(g['gdepth'].norm()) < 1e-3
                else:
                    assert rel(depth.grad, g['gdepth']) < 1e-3, (pname, rot, pad)
                assert rel(p.grad, g['gpose']) < 1e-4, (pname, rot, pad)


def test_g4_photometric(golden):
    g4 = golden('g4_photometric')
    B, H, W = 2, 64, 96
    for R, use_mask in ((2, False), (4, True)):
        tgt = I.images(B, H, W, seed=60)
        refs = [I.images(B, H, W, seed=61 + r) for r in range(R)]
        K, Kinv = I.intrinsics(B, H / 128.0)
        for rot, pad in (('euler', 'zeros'), ('quat', 'border')):
            depth = [I.depth_map(B, H >> s, W >> s, seed=70 + s).unsqueeze(1).requires_grad_(True) for s in range(4)]
            pose = I.poses(B, R, seed=80).requires_grad_(True)
            masks = [I.mask_map(B, R, H >> s, W >> s, seed=90 + s).requires_grad_(True) for s in range(4)] \
                if use_mask else [None] * 4
            g = g4['R%d_%s_%s' % (R, rot, pad)]
            loss = OL.photometric_reconstruction_loss(tgt, refs, K, Kinv, depth, masks, pose, rot, pad)
            assert abs(float(loss) - float(g['loss'])) < 1e-4 * abs(float(g['loss']))
            loss.backward()
            for d, r in zip(depth, g['gdepth']):
                assert rel(d.grad, r) < 1e-3
            assert rel(pose.grad, g['gpose']) < 1e-3
            if use_mask:
                for m, r in zip(masks, g['gmask']):
                    assert rel(m.grad, r) < 1e-3
    masks = [I.mask_map(B, 4, H >> s, W >> s, seed=90 + s).requires_grad_(True) for s in range(4)]
    le = OL.explainability_loss(masks)
    assert abs(float(le) - float(g4['explainability']['loss'])) < 1e-5
    le.backward()
    for m, r in zip(masks, g4['explainability']['gmask']):
        assert rel(m.grad, r) < 1e-5


def test_g5_smooth(golden):
    g5 = golden('g5_smooth')
    ramp = torch.arange(52.).view(1, 1, 1, 52).expand(2, 1, 16, 52).contiguous()
    assert float(OL.smooth_loss([ramp])) == 0.0 == float(g5['ramp'])
    assert float(OL.smooth_loss([ramp * ramp])) == pytest.approx(2.0, abs=1e-6)
    assert float(g5['x2']) == pytest.approx(2.0, abs=1e-6)
    maps = [I.depth_map(2, 64 >> s, 96 >> s, seed=100 + s).unsqueeze(1).requires_grad_(True) for s in range(4)]
    l = OL.smooth_loss(maps)
    assert float(l) == pytest.approx(float(g5['random']['loss']), rel=1e-5)
    l.backward()
    for m, r in zip(maps, g5['random']['grads']):
        assert rel(m.grad, r) < 1e-5


@pytest.mark.parametrize('ds', ['kitti', 'nyu'])
def test_g6_l1(golden, ds):
    g6 = golden('g6_l1')
    gt = I.sparse_gt(3, 64, 96, seed=110, dataset=ds)
    pred = I.depth_map(3, 64, 96, seed=111, lo=0.0005, hi=95.0 if ds == 'kitti' else 12.0).unsqueeze(1).requires_grad_(True)
    l = OL.l1_loss(gt, [pred], ds)
    assert float(l) == pytest.approx(float(g6[ds]['loss']), rel=1e-5)
    l.backward()
    assert rel(pred.grad, g6[ds]['grad']) < 1e-5
    gt2 = gt.clone()
    gt2[1] = 0
    assert math.isnan(float(OL.l1_loss(gt2, [pred.detach()], ds))) and math.isnan(float(g6[ds + '_empty']))


def test_g7_compute_errors(golden):
    g7 = golden('g7_errors')
    gt = I.sparse_gt(3, 128, 416, seed=120, dataset='kitti', density=0.2)
    pred = I.depth_map(3, 128, 416, seed=121, lo=0.0005, hi=95.0)
    for a, b in zip(OL.compute_errors(gt, pred, 'kitti', True), g7['kitti_crop']):
        assert a == pytest.approx(b, rel=1e-5)
    for a, b in zip(OL.compute_errors(gt, pred, 'kitti', True, True), g7['kitti_crop_unsup']):
        assert a == pytest.approx(b, rel=1e-5)
    assert OL.garg_crop(128, 416) == (52, 126, 14, 401)
    assert OL.error_counters(gt, pred, 'kitti', True).tolist() == g7['kitti_crop_counters']      # bit-exact ints
    gtn = I.sparse_gt(2, 64, 96, seed=122, dataset='nyu', density=0.9)
    predn = I.depth_map(2, 64, 96, seed=123, lo=0.0005, hi=12.0)
    for a, b in zip(OL.compute_errors(gtn, predn, 'nyu', False), g7['nyu']):
        assert a == pytest.approx(b, rel=1e-5)


def test_g9_layers_terms(golden):
    g9 = golden('g9_layers')
    x = I.images(2, 32, 48, seed=130) * 0.5 + 0.5
    y = I.images(2, 32, 48, seed=131) * 0.5 + 0.5
    assert rel(OL.ssim(x, y), g9['ssim']) < 1e-5
    disp = I.depth_map(2, 32, 48, seed=132).unsqueeze(1)
    assert float(OL.get_smooth_loss(disp, x)) == pytest.approx(float(g9['edge_smooth']), rel=1e-5)
    a = I.depth_map(1, 8, 200, seed=133).flatten()
    b = I.depth_map(1, 8, 200, seed=134).flatten()
    for u, v in zip(OL.compute_depth_errors(a, b), g9['depth_errors']):
        assert float(u) == pytest.approx(v, rel=1e-5)


def test_g8_train_trajectory_oracle_port(golden):
    """G8: three Adam steps of the unmodified reference `train.train` (train.py:394-539; fixture made by
    oracle/make_golden.py:make_g8) against the same three steps on the oracle port -- pins the optimizer-coupled behaviour
    (BatchNorm running statistics after more than one step, Adam on rounding-noise gradients) of the restatement."""
    from oracle.make_golden import G8_KEYS, g8_batches
    g = golden('g8_train_trajectory')
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    params = [v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k]
    opt = torch.optim.Adam(params, lr=2e-4, betas=(0.9, 0.999))
    rows = []
    for x, gt in g8_batches():
        disp = ON.disp_vgg_bn(sd, x, True)
        depth = [1 / d for d in disp]
        l1, l3 = OL.l1_loss(gt, depth, 'kitti'), OL.smooth_loss(depth)
        loss = 1.0 * l1 + 0.0 * l3
        opt.zero_grad()
        loss.backward()
        opt.step()
        rows.append([float(loss), float(l1), 0.0, float(l3)])
    for r, q in zip(rows, g['rows']):
        for a, b in zip(r, q):
            assert abs(a - b) <= 2e-4 * max(abs(b), 1e-6), (rows, g['rows'])
    for k in G8_KEYS:
        # parameter UPDATES (Adam moves every element by ~lr per step whatever the gradient's size, so elements whose
        # gradient is round-off noise differ in sign between two fp32 implementations: compare on a loose L2 scale)
        du, dr = I.subsample(sd[k].detach()) - g['init'][k], g['params'][k] - g['init'][k]
        assert float((du - dr).norm() / dr.norm().clamp_min(1e-30)) < 0.2, k
    for k, v in g['running'].items():
        # the conv bias in front of this BatchNorm has an analytically zero gradient; Adam turns its round-off noise into
        # +-lr steps per iteration, which shift the batch mean (not the variance) by the same amount on either side
        if k.endswith('running_mean'):
            assert float((sd[k] - v).abs().max()) < 3 * 2e-4, k
        else:
            assert rel(sd[k].float(), v.float()) < 1e-4, k


G10_CASES = [('l2_kitti', 'kitti', 95.0), ('l2_nyu', 'nyu', 12.0), ('berhu_kitti', 'kitti', 95.0), ('scale_inv_kitti', 'kitti', 95.0),
             ('scale_inv_nyu', 'nyu', 12.0), ('multi_l1', 'kitti', 95.0), ('multi_l1_max', 'kitti', 95.0), ('multi_full_l1', 'kitti', 95.0),
             ('multi_l2', 'kitti', 95.0), ('multi_berhu', 'kitti', 95.0), ('multi_scale_inv', 'kitti', 95.0)]


def g10_inputs(ds, hi, device='cpu'):
    B, H, W = 3, 32, 64
    gt = I.sparse_gt(B, H, W, seed=140 if ds == 'kitti' else 141, dataset=ds, density=0.3 if ds == 'kitti' else 0.9)
    preds = [I.depth_map(B, H >> s, W >> s, seed=150 + s, lo=0.0005, hi=hi).unsqueeze(1).to(device).requires_grad_(True) for s in range(4)]
    return gt.to(device), preds


def g10_call(mod, name, gt, preds, ds, oracle):
    """Dispatch one G10 case on the oracle (oracle=True) or on a module with the reference's function names."""
    if oracle:
        return {'l2': lambda: mod.l2_loss(gt, preds, ds), 'berhu': lambda: mod.berhu_loss(gt, preds, ds),
                'scale_inv': lambda: mod.scale_invariant_loss(gt, preds, ds), 'multi_l1': lambda: mod.multiscale_loss('l1', gt, preds),
                'multi_l1_max': lambda: mod.multiscale_loss('l1', gt, preds, 'max'),
                'multi_full_l1': lambda: mod.multiscale_full_l1_loss(gt, preds), 'multi_l2': lambda: mod.multiscale_loss('l2', gt, preds),
                'multi_berhu': lambda: mod.multiscale_loss('berhu', gt, preds),
                'multi_scale_inv': lambda: mod.multiscale_loss('scale_inv', gt, preds)}[name.replace('_kitti', '').replace('_nyu', '')]()
    return {'l2': lambda: mod.l2_loss(gt, preds, ds), 'berhu': lambda: mod.berhu_loss(gt, preds, ds),
            'scale_inv': lambda: mod.Scale_invariant_loss(gt, preds, ds), 'multi_l1': lambda: mod.Multiscale_L1_loss(gt, preds),
            'multi_l1_max': lambda: mod.Multiscale_L1_loss(gt, preds, 'max'),
            'multi_full_l1': lambda: mod.Multiscale_FULL_L1_loss(gt, preds), 'multi_l2': lambda: mod.Multiscale_L2_loss(gt, preds),
            'multi_berhu': lambda: mod.Multiscale_berhu_loss(gt, preds),
            'multi_scale_inv': lambda: mod.Multiscale_scale_inv_loss(gt, preds)}[name.replace('_kitti', '').replace('_nyu', '')]()


@pytest.mark.parametrize('name,ds,hi', G10_CASES)
def test_g10_supervised_losses(golden, name, ds, hi):
    """The remaining `--loss` choices of train.py:449-470 (loss_functions.py:77-315): oracle restatement vs the reference."""
    g = golden('g10_supervised_losses')[name]
    gt, preds = g10_inputs(ds, hi)
    l = g10_call(OL, name, gt, preds, ds, True)
    l.backward()
    assert abs(float(l) - float(g['loss'])) <= 2e-5 * abs(float(g['loss'])), (float(l), float(g['loss']))
    for p, r in zip(preds, g['grads']):
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            assert rel(p.grad, r) < 2e-5


@pytest.mark.parametrize('tag', ['vgg_kitti', 'res50_nyu'])
def test_g11_validate_with_gt_oracle(golden, tag):
    """validate_with_gt (train.py:642-723) restated on the oracle vs the reference's own run (fixture G11)."""
    from oracle.make_golden import g11_loaders
    g = golden('g11_validate')[tag]
    kitti, nyu = g11_loaders()
    if tag == 'vgg_kitti':
        sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
        sd.update({k: v.clone() for k, v in g['running'].items()})
        errs = OL.validate_with_gt(lambda x: ON.disp_vgg_bn(sd, x, False, 'kitti'), kitti, 'kitti')
    else:
        sd = ON.init_state_dict('Disp_res_50', 0)
        sd.update({k: v.clone() for k, v in g['running'].items()})
        errs = OL.validate_with_gt(lambda x: ON.disp_res_50(sd, x, False, 'nyu'), nyu, 'nyu')
    for a, b in zip(errs, g['errors']):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), (errs, g['errors'])
