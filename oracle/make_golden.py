"""Golden-vector generator (TEST INFRASTRUCTURE ONLY; runs in the authoring container only).

Imports the *unmodified* reference from /root/reference (read-only), runs it on CPU in fp32 on seeded
synthetic inputs and writes small fixtures to tests/golden/*.pt.  The reference cannot travel to the GPU
box, the fixtures can.  Inputs are regenerated from their seed by tests/_inputs.py (shared by this script
and the tests), network weights by oracle.nets.init_state_dict (which must reproduce
``torch.manual_seed(s); net.init_weights()`` bit for bit -- fixture G0 checks exactly that via checksums).

Run:  python -m oracle.make_golden            (from the repo root)
"""
import os
import sys
import types
import warnings

import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
sys.path.insert(0, os.path.join(os.path.dirname(OUT)))            # tests/ for _inputs
warnings.filterwarnings('ignore')


def import_reference():
    """Shims of SURVEY.md section 8(c): scipy.misc.imresize, and the HF `datasets` name clash."""
    import scipy.misc
    if not hasattr(scipy.misc, 'imresize'):
        scipy.misc.imresize = lambda *a, **k: None                 # imported, never called on the hot path
    sys.path.insert(0, REF)
    for m in ('models', 'loss_functions', 'inverse_warp', 'layers'):
        sys.modules.pop(m, None)
    import models as ref_models
    import loss_functions as ref_loss
    import inverse_warp as ref_warp
    import layers as ref_layers
    return ref_models, ref_loss, ref_warp, ref_layers


def checksum(sd):
    """Order-independent fingerprint of a state_dict: per-key (sum, abs-sum) in float64."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()
            if v.dtype.is_floating_point}


def main():
    import _inputs as I
    os.makedirs(OUT, exist_ok=True)
    rm, rl, rw, rlay = import_reference()
    torch.set_grad_enabled(True)
    G = {}

    # ---- G0: weight-init fingerprints (pins oracle.nets.init_state_dict to the reference's init) ------------
    fp = {}
    for name, ctor in (('DispNetS', lambda: rm.DispNetS()), ('Disp_vgg_BN', lambda: rm.Disp_vgg_BN()),
                       ('Disp_res_50', lambda: rm.Disp_res_50()),
                       ('PoseExpNet_r2', lambda: rm.PoseExpNet(2, False)),
                       ('PoseExpNet_r4e', lambda: rm.PoseExpNet(4, True))):
        net = ctor()
        torch.manual_seed(0)
        net.init_weights()
        fp[name] = checksum(net.state_dict())
    torch.save(fp, os.path.join(OUT, 'g0_init_fingerprints.pt'))

    def seeded(ctor, seed=0):
        net = ctor()
        torch.manual_seed(seed)
        net.init_weights()
        return net

    # ---- G1: DispNetS eval forward, BASELINE config 1 ------------------------------------------------------
    net = seeded(lambda: rm.DispNetS()).eval()
    with torch.no_grad():
        d = net(I.images(1, 128, 416, seed=1))
    G['g1_dispnets_eval'] = d.clone()
    # DispNetS train-mode 4 scales + grads of a scalar at reduced size (exercises crop_like at 1x4 bottleneck)
    net.train()
    x = I.images(2, 128, 160, seed=2)
    outs = net(x)
    loss = sum((o * I.probe_like(o, 10 + i)).sum() for i, o in enumerate(outs))
    loss.backward()
    G['g1_dispnets_train'] = dict(outs=[o.detach().clone() for o in outs],
                                  grads={k: I.subsample(p.grad) for k, p in net.named_parameters()
                                         if k in ('conv1.0.weight', 'conv7.2.bias', 'upconv7.0.weight',
                                                  'iconv3.0.weight', 'predict_disp1.0.weight')})

    # ---- G2: Disp_vgg_BN train-mode forward/backward + BN running stats after one step, b=2 ---------------
    net = seeded(lambda: rm.Disp_vgg_BN()).train()
    x = I.images(2, 64, 96, seed=3)
    outs = net(x)
    loss = sum((o * I.probe_like(o, 20 + i)).sum() for i, o in enumerate(outs))
    loss.backward()
    sd = net.state_dict()
    G['g2_vgg_train'] = dict(
        outs=[o.detach().clone() for o in outs],
        running={k: sd[k].clone() for k in sd if 'running' in k or 'num_batches' in k},
        grads={k: I.subsample(p.grad) for k, p in net.named_parameters()
               if p.grad is not None and k in ('features.features.0.weight', 'features.features.1.weight',
                                               'features.features.1.bias', 'features.features.40.weight',
                                               'features.features.41.weight', 'upconv4.0.weight',
                                               'iconv2.0.weight', 'iconv0.0.bias', 'disp0.0.weight',
                                               'disp3.0.bias')},
        no_grad_keys=[k for k, p in net.named_parameters() if p.grad is None])
    net = seeded(lambda: rm.Disp_vgg_BN()).eval()      # fresh running stats (0 / 1)
    with torch.no_grad():
        G['g2_vgg_eval'] = net(I.images(1, 128, 416, seed=4)).clone()

    # ---- Disp_res_50 and PoseExpNet -------------------------------------------------------------------------
    net = seeded(lambda: rm.Disp_res_50()).train()
    x = I.images(2, 64, 96, seed=5)
    outs = net(x)
    loss = sum((o * I.probe_like(o, 30 + i)).sum() for i, o in enumerate(outs))
    loss.backward()
    sd = net.state_dict()
    G['g2b_res50_train'] = dict(
        outs=[o.detach().clone() for o in outs],
        running={k: sd[k].clone() for k in ('bn1.running_mean', 'bn1.running_var', 'layer1.0.bn1.running_mean',
                                            'layer4.2.bn3.running_var', 'layer2.0.downsample.1.running_mean')},
        grads={k: I.subsample(p.grad) for k, p in net.named_parameters()
               if p.grad is not None and k in ('conv1.weight', 'layer1.0.conv1.weight', 'layer3.0.downsample.0.weight',
                                               'layer4.2.bn3.weight', 'upconv5.0.weight', 'iconv1.0.weight')},
        no_grad_keys=[k for k, p in net.named_parameters() if p.grad is None])
    for tag, R, exp in (('r2', 2, False), ('r4e', 4, True)):
        net = seeded(lambda: rm.PoseExpNet(R, exp)).train()
        hw = (128, 416) if not exp else (64, 128)
        tgt = I.images(2, hw[0], hw[1], seed=6)
        refs = [I.images(2, hw[0], hw[1], seed=7 + r) for r in range(R)]
        masks, pose = net(tgt, refs)
        loss = (pose * I.probe_like(pose, 40)).sum()
        if exp:
            loss = loss + sum((m * I.probe_like(m, 41 + i)).sum() for i, m in enumerate(masks))
        loss.backward()
        G['g2c_pose_' + tag] = dict(pose=pose.detach().clone(),
                                    masks=[None if m is None else m.detach().clone() for m in masks],
                                    grads={k: I.subsample(p.grad) for k, p in net.named_parameters()
                                           if k in ('conv1.0.weight', 'pose_pred.bias', 'conv7.0.weight',
                                                    'upconv5.0.weight', 'predict_mask1.weight')})

    # ---- G3: inverse_warp identity / random pose, euler & quat, zeros & border -----------------------------
    g3 = {}
    B, h, w = 2, 32, 104
    img = I.images(B, h, w, seed=50)
    K, Kinv = I.intrinsics(B, h / 128.0)
    for pname, pose in (('identity', torch.zeros(B, 6)), ('random', I.poses(B, 1, seed=51)[:, 0])):
        for rot in ('euler', 'quat'):
            for pad in ('zeros', 'border'):
                depth = I.depth_map(B, h, w, seed=52).requires_grad_(True)
                p = pose.clone().requires_grad_(True)
                out = rw.inverse_warp(img, depth, p, K, Kinv, rot, pad)
                (out * I.probe_like(out, 53)).sum().backward()
                g3['%s_%s_%s' % (pname, rot, pad)] = dict(out=out.detach().clone(), gdepth=depth.grad.clone(),
                                                          gpose=p.grad.clone())
    G['g3_inverse_warp'] = g3

    # ---- G4: photometric loss + grads, R in {2,4}, masks on/off -------------------------------------------
    g4 = {}
    B, H, W = 2, 64, 96
    for R, use_mask in ((2, False), (4, True)):
        tgt = I.images(B, H, W, seed=60)
        refs = [I.images(B, H, W, seed=61 + r) for r in range(R)]
        K, Kinv = I.intrinsics(B, H / 128.0)
        depth = [I.depth_map(B, H >> s, W >> s, seed=70 + s).unsqueeze(1).requires_grad_(True) for s in range(4)]
        pose = I.poses(B, R, seed=80).requires_grad_(True)
        masks = [I.mask_map(B, R, H >> s, W >> s, seed=90 + s).requires_grad_(True) for s in range(4)] \
            if use_mask else [None] * 4
        for rot, pad in (('euler', 'zeros'), ('quat', 'border')):
            for t in depth + [pose] + [m for m in masks if m is not None]:
                t.grad = None
            loss = rl.photometric_reconstruction_loss(tgt, refs, K, Kinv, depth, masks, pose, rot, pad)
            loss.backward()
            g4['R%d_%s_%s' % (R, rot, pad)] = dict(
                loss=loss.detach().clone(), gdepth=[d.grad.clone() for d in depth], gpose=pose.grad.clone(),
                gmask=[m.grad.clone() for m in masks] if use_mask else None)
        if use_mask:
            for m in masks:
                m.grad = None
            le = rl.explainability_loss(masks)
            le.backward()
            g4['explainability'] = dict(loss=le.detach().clone(), gmask=[m.grad.clone() for m in masks])
    G['g4_photometric'] = g4

    # ---- G5: smooth_loss KATs + random ---------------------------------------------------------------------
    g5 = {}
    ramp = torch.arange(52.).view(1, 1, 1, 52).expand(2, 1, 16, 52).contiguous()
    g5['ramp'] = rl.smooth_loss([ramp]).clone()
    g5['x2'] = rl.smooth_loss([ramp * ramp]).clone()
    maps = [I.depth_map(2, 64 >> s, 96 >> s, seed=100 + s).unsqueeze(1).requires_grad_(True) for s in range(4)]
    ls = rl.smooth_loss(maps)
    ls.backward()
    g5['random'] = dict(loss=ls.detach().clone(), grads=[m.grad.clone() for m in maps])
    G['g5_smooth'] = g5

    # ---- G6: l1_loss with sparse gt incl. an all-invalid sample -------------------------------------------
    g6 = {}
    for ds in ('kitti', 'nyu'):
        gt = I.sparse_gt(3, 64, 96, seed=110, dataset=ds)
        pred = I.depth_map(3, 64, 96, seed=111, lo=0.0005, hi=95.0 if ds == 'kitti' else 12.0).unsqueeze(1).requires_grad_(True)
        l = rl.l1_loss(gt, [pred], ds)
        l.backward()
        g6[ds] = dict(loss=l.detach().clone(), grad=pred.grad.clone())
        gt2 = gt.clone()
        gt2[1] = 0                                  # all-invalid sample -> NaN (hard part 7)
        g6[ds + '_empty'] = rl.l1_loss(gt2, [pred.detach()], ds).clone()
    G['g6_l1'] = g6

    # ---- G7: compute_errors incl. crop window and the oracle's own int counters ----------------------------
    g7 = {}
    gt = I.sparse_gt(3, 128, 416, seed=120, dataset='kitti', density=0.2)
    pred = I.depth_map(3, 128, 416, seed=121, lo=0.0005, hi=95.0)
    g7['kitti_crop'] = rl.compute_errors(gt, pred, 'kitti', True)
    g7['kitti_crop_unsup'] = rl.compute_errors(gt, pred, 'kitti', True, True)
    gtn = I.sparse_gt(2, 64, 96, seed=122, dataset='nyu', density=0.9)
    predn = I.depth_map(2, 64, 96, seed=123, lo=0.0005, hi=12.0)
    g7['nyu'] = rl.compute_errors(gtn, predn, 'nyu', False)
    # reference-derived integer counters: count/n_valid per sample recomputed with the reference's own ops
    cnt = []
    cm = torch.zeros(128, 416, dtype=torch.bool)
    cm[int(0.40810811 * 128):int(0.99189189 * 128), int(0.03594771 * 416):int(0.96405229 * 416)] = True
    for g, p in zip(gt, pred):
        valid = (g > 0) & (g < 80) & cm
        vg, vp = g[valid], p[valid].clamp(1e-3, 80)
        th = torch.max(vg / vp, vp / vg)
        cnt.append([int(valid.sum()), int((th < 1.25).sum()), int((th < 1.25 ** 2).sum()), int((th < 1.25 ** 3).sum())])
    g7['kitti_crop_counters'] = cnt
    G['g7_errors'] = g7

    # ---- layers.py optional terms ---------------------------------------------------------------------------
    g9 = {}
    x = I.images(2, 32, 48, seed=130) * 0.5 + 0.5
    y = I.images(2, 32, 48, seed=131) * 0.5 + 0.5
    g9['ssim'] = rlay.SSIM()(x, y).clone()
    disp = I.depth_map(2, 32, 48, seed=132).unsqueeze(1)
    g9['edge_smooth'] = rlay.get_smooth_loss(disp, x).clone()
    a = I.depth_map(1, 8, 200, seed=133).flatten()
    b = I.depth_map(1, 8, 200, seed=134).flatten()
    g9['depth_errors'] = [float(v) for v in rlay.compute_depth_errors(a, b)]
    G['g9_layers'] = g9

    for k, v in G.items():
        torch.save(v, os.path.join(OUT, k + '.pt'))
        print('%-24s %8.1f KB' % (k, os.path.getsize(os.path.join(OUT, k + '.pt')) / 1024))


G8_KEYS = ('features.features.0.weight', 'features.features.1.weight', 'features.features.40.weight', 'upconv4.0.weight',
           'iconv2.0.weight', 'iconv0.0.weight', 'disp0.0.weight', 'disp3.0.bias')


def g8_batches(n=3, b=4, h=128, w=416):
    import _inputs as I
    return [(I.images(b, h, w, seed=300 + i), I.sparse_gt(b, h, w, seed=310 + i, dataset='kitti')) for i in range(n)]


def make_g8():
    """G8: three steps of the UNMODIFIED reference `train.train` (train.py:394-539) with Adam on Disp_vgg_BN,
    b=4, 128x416, `--loss L1`: the per-step [loss, loss_1, loss_2, loss_3] rows the loop itself logs (train.py:530-532), the
    returned average, and a few parameters / BatchNorm buffers after the third optimizer step."""
    import csv
    import tempfile
    from oracle import refshim as R
    ref = R.import_reference(REF)
    T = ref.train
    net = ref.models.Disp_vgg_BN()
    torch.manual_seed(0)               # same order as fixture G0: construct, seed, init_weights()
    net.init_weights()
    init = {k: v.clone() for k, v in net.state_dict().items()}
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-4, betas=(0.9, 0.999), weight_decay=0)
    with tempfile.TemporaryDirectory() as d:
        args = R.reference_args(d, batch_size=4)
        T.device, T.n_iter = torch.device('cpu'), 0
        avg = T.train(args, g8_batches(), net, torch.nn.Identity(), opt, 3, R.NullLogger(), R.NullWriter())
        rows = [[float(v) for v in r] for r in csv.reader(open(os.path.join(d, args.log_full)), delimiter='\t')]
    sd = net.state_dict()
    out = dict(rows=rows, avg=float(avg),
               params={k: I_sub(sd[k]) for k in G8_KEYS}, init={k: I_sub(init[k]) for k in G8_KEYS},
               running={k: sd[k].clone() for k in sd if k.startswith('features.features.1.') and ('running' in k or 'num_batches' in k)})
    torch.save(out, os.path.join(OUT, 'g8_train_trajectory.pt'))
    print('g8 rows', rows, 'avg', avg)


def make_g10():
    """G10: the supervised losses behind train.py's --loss switch other than L1 (loss_functions.py:77-315), values and
    gradients w.r.t. every scale's prediction, from the unmodified reference."""
    import _inputs as I
    rm, rl, rw, rlay = import_reference()
    B, H, W = 3, 32, 64
    gt = I.sparse_gt(B, H, W, seed=140, dataset='kitti', density=0.3)
    gtn = I.sparse_gt(B, H, W, seed=141, dataset='nyu', density=0.9)
    out = {}

    def preds(hi):
        return [I.depth_map(B, H >> s, W >> s, seed=150 + s, lo=0.0005, hi=hi).unsqueeze(1).requires_grad_(True) for s in range(4)]
    cases = [('l2_kitti', lambda d: rl.l2_loss(gt, d, 'kitti'), 95.0), ('l2_nyu', lambda d: rl.l2_loss(gtn, d, 'nyu'), 12.0),
             ('berhu_kitti', lambda d: rl.berhu_loss(gt, d, 'kitti'), 95.0),
             ('scale_inv_kitti', lambda d: rl.Scale_invariant_loss(gt, d, 'kitti'), 95.0),
             ('scale_inv_nyu', lambda d: rl.Scale_invariant_loss(gtn, d, 'nyu'), 12.0),
             ('multi_l1', lambda d: rl.Multiscale_L1_loss(gt, d), 95.0),
             ('multi_l1_max', lambda d: rl.Multiscale_L1_loss(gt, d, 'max'), 95.0),
             ('multi_full_l1', lambda d: rl.Multiscale_FULL_L1_loss(gt, d), 95.0),
             ('multi_l2', lambda d: rl.Multiscale_L2_loss(gt, d), 95.0),
             ('multi_berhu', lambda d: rl.Multiscale_berhu_loss(gt, d), 95.0),
             ('multi_scale_inv', lambda d: rl.Multiscale_scale_inv_loss(gt, d), 95.0)]
    for name, fn, hi in cases:
        d = preds(hi)
        l = fn(d)
        l.backward()
        out[name] = dict(loss=l.detach().clone(), grads=[None if t.grad is None else t.grad.clone() for t in d])
        print('%-18s %.6f' % (name, float(l)))
    try:
        rl.berhu_loss(gtn, preds(12.0), 'nyu')
        out['berhu_nyu_raises'] = None
    except Exception as e:  # noqa: BLE001
        out['berhu_nyu_raises'] = type(e).__name__
    torch.save(out, os.path.join(OUT, 'g10_supervised_losses.pt'))


def g11_loaders():
    import _inputs as I
    kitti = [(I.images(2, 128, 416, seed=600 + i), I.sparse_gt(2, 128, 416, seed=610 + i, dataset='kitti', density=0.05)) for i in range(2)]
    # NYU: the loader yields [B, 2, H, W] (depth, mask); the ground truth is larger than the network output (train.py:696-700)
    nyu = [(I.images(2, 128, 160, seed=620 + i),
            torch.stack([I.sparse_gt(2, 150, 200, seed=630 + i, dataset='nyu', density=0.9), torch.ones(2, 150, 200)], 1)) for i in range(2)]
    return kitti, nyu


def make_g11():
    """G11: the reference's validate_with_gt (train.py:642-723) on two synthetic validation batches: Disp_vgg_BN / kitti
    (Garg crop) and Disp_res_50 / nyu (prediction up-sampled to the ground truth's resolution), BatchNorm running
    statistics taken after one training step so that eval mode is not the identity normalisation."""
    from oracle import refshim as R
    ref = R.import_reference(REF)
    T = ref.train
    T.device = torch.device('cpu')
    kitti, nyu = g11_loaders()
    out = {}
    for tag, ctor, loader, ds in (('vgg_kitti', lambda: ref.models.Disp_vgg_BN('kitti'), kitti, 'kitti'),
                                  ('res50_nyu', lambda: ref.models.Disp_res_50('nyu'), nyu, 'nyu')):
        net = ctor()
        torch.manual_seed(0)
        net.init_weights()
        net.train()
        with torch.no_grad():
            net(loader[0][0])                       # one training-mode forward: running statistics move off (0, 1)
        sd = {k: v.clone() for k, v in net.state_dict().items() if 'running' in k or 'num_batches' in k}
        args = R.reference_args('/tmp', dataset=ds)
        errs, names = T.validate_with_gt(args, loader, net, 0, R.NullLogger(), [])
        out[tag] = dict(errors=[float(e) for e in errs], names=names, running=sd)
        print(tag, [round(float(e), 5) for e in errs])
    torch.save(out, os.path.join(OUT, 'g11_validate.pt'))


def I_sub(t):
    import _inputs as I
    return I.subsample(t)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'g8':
        make_g8()
    elif len(sys.argv) > 1 and sys.argv[1] == 'g10':
        make_g10()
        make_g11()
    elif len(sys.argv) > 1 and sys.argv[1] == 'g11':
        make_g11()
    else:
        main()
        make_g8()
        make_g10()
        make_g11()
