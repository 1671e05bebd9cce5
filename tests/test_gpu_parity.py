"""GPU parity tests proper: the CUDA product path (through the C ABI of libdispnet_b200.so) against

  (a) the committed golden fixtures that oracle/make_golden.py generated from the reference itself, and
  (b) the CPU oracle on the same seeded inputs.

Tolerances (north star: 1e-3 relative fp32 on disparity maps and loss scalars; integer counters bit-exact):
  * loss / geometry kernels compute in fp32: loss scalars 1e-5 relative, gradients 1e-3 (L2-relative; fp32 warps amplify
    coordinate round-off, the oracle itself differs from the reference by ~1e-4 there);
  * precision 'fp32' (CUDA-core kernels): disparities 1e-5; parameter gradients within 3e-3 of the fp32 oracle, which is
    the oracle's own fp32-vs-fp64 noise floor on these BatchNorm stacks (measured: VGG 1.6e-3);
  * precision 'mixed' (tcgen05: fp16 operands, fp32 accumulate, bf16 gradient activations): single layers 1e-3; whole
    networks 3e-3 on disparities (18 stacked fp16 roundings; measured 1.5e-3 at 128x416) and 1e-3 on the loss scalar;
    gradients are checked for direction (cosine > 0.99), as bf16 gradient storage is a precision choice of the product.
"""
import math

import pytest
import torch

import _inputs as I

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), 'these tests need a CUDA device (run with -m gpu on the B200 box)'
    from supervised_dispnet_b200 import _lib as L
    assert L.lib().dn_version() >= 100
    yield


def P():
    import _parity
    return _parity


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_robust(a, b, k=2):
    """L2-relative error after discarding the k worst elements: a warp coordinate that lands within one ulp of a
    clipping / in-view boundary takes the other branch under a different fp32 operation order (SURVEY.md hard part 5)
    and changes exactly that pixel's gradient; everything else must agree."""
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    d = (a - b).abs()
    keep = torch.ones_like(d, dtype=torch.bool)
    keep[d.topk(k).indices] = False
    return float((a[keep] - b[keep]).norm() / b[keep].norm().clamp_min(1e-30))


DEV = 'cuda'

# ------------------------------------------------------------------------------------------------------------
# (a) golden fixtures generated from the reference
# ------------------------------------------------------------------------------------------------------------


def test_golden_g3_inverse_warp(golden):
    from supervised_dispnet_b200.inverse_warp import inverse_warp
    g3 = golden('g3_inverse_warp')
    B, h, w = 2, 32, 104
    img = I.images(B, h, w, seed=50).to(DEV)
    K, Kinv = [t.to(DEV) for t in I.intrinsics(B, h / 128.0)]
    for pname, pose in (('identity', torch.zeros(B, 6)), ('random', I.poses(B, 1, seed=51)[:, 0])):
        for rot in ('euler', 'quat'):
            for pad in ('zeros', 'border'):
                g = g3['%s_%s_%s' % (pname, rot, pad)]
                depth = I.depth_map(B, h, w, seed=52).to(DEV).requires_grad_(True)
                p = pose.clone().to(DEV).requires_grad_(True)
                out = inverse_warp(img, depth, p, K, Kinv, rot, pad)
                bad = ((out.cpu() - g['out']).abs() > 1e-4).float().mean().item()
                assert bad < 2e-3, (pname, rot, pad, bad)
                (out * I.probe_like(out, 53).to(DEV)).sum().backward()
                if pname == 'identity':
                    assert float(depth.grad.norm()) < 1e-3
                else:
                    assert rel(depth.grad, g['gdepth']) < 1e-3, (pname, rot, pad)
                assert rel(p.grad, g['gpose']) < 1e-3, (pname, rot, pad)


def test_golden_g4_photometric(golden):
    from supervised_dispnet_b200 import loss_functions as LF
    g4 = golden('g4_photometric')
    B, H, W = 2, 64, 96
    for R, use_mask in ((2, False), (4, True)):
        tgt = I.images(B, H, W, seed=60).to(DEV)
        refs = [I.images(B, H, W, seed=61 + r).to(DEV) for r in range(R)]
        K, Kinv = [t.to(DEV) for t in I.intrinsics(B, H / 128.0)]
        for rot, pad in (('euler', 'zeros'), ('quat', 'border')):
            depth = [I.depth_map(B, H >> s, W >> s, seed=70 + s).unsqueeze(1).to(DEV).requires_grad_(True) for s in range(4)]
            pose = I.poses(B, R, seed=80).to(DEV).requires_grad_(True)
            masks = [I.mask_map(B, R, H >> s, W >> s, seed=90 + s).to(DEV).requires_grad_(True) for s in range(4)] \
                if use_mask else [None] * 4
            g = g4['R%d_%s_%s' % (R, rot, pad)]
            loss = LF.photometric_reconstruction_loss(tgt, refs, K, Kinv, depth, masks, pose, rot, pad)
            assert abs(float(loss) - float(g['loss'])) < 1e-4 * abs(float(g['loss']))
            loss.backward()
            for d, r in zip(depth, g['gdepth']):
                assert rel_robust(d.grad, r) < 1e-3
                assert ((d.grad.cpu() - r).abs() > 1e-3 * r.abs().max()).sum() <= 2
            # a sample that flips in / out of view on a 1-ulp coordinate difference moves one pose component by ~1e-3
            assert rel(pose.grad, g['gpose']) < (1e-2 if pad == 'border' else 1e-3)   # border: the 1-2 flipped pixels above
            if use_mask:
                for m, r in zip(masks, g['gmask']):
                    assert rel(m.grad, r) < 1e-3
    masks = [I.mask_map(B, 4, H >> s, W >> s, seed=90 + s).to(DEV).requires_grad_(True) for s in range(4)]
    le = LF.explainability_loss(masks)
    assert abs(float(le) - float(g4['explainability']['loss'])) < 1e-5
    le.backward()
    for m, r in zip(masks, g4['explainability']['gmask']):
        assert rel(m.grad, r) < 1e-5


def test_golden_g5_smooth(golden):
    from supervised_dispnet_b200 import loss_functions as LF
    g5 = golden('g5_smooth')
    ramp = torch.arange(52.).view(1, 1, 1, 52).expand(2, 1, 16, 52).contiguous().to(DEV)
    assert float(LF.smooth_loss([ramp])) == 0.0 == float(g5['ramp'])
    assert float(LF.smooth_loss([ramp * ramp])) == pytest.approx(2.0, abs=1e-6)
    maps = [I.depth_map(2, 64 >> s, 96 >> s, seed=100 + s).unsqueeze(1).to(DEV).requires_grad_(True) for s in range(4)]
    l = LF.smooth_loss(maps)
    assert float(l) == pytest.approx(float(g5['random']['loss']), rel=1e-5)
    l.backward()
    for m, r in zip(maps, g5['random']['grads']):
        assert rel(m.grad, r) < 1e-5


@pytest.mark.parametrize('ds', ['kitti', 'nyu'])
def test_golden_g6_l1(golden, ds):
    from supervised_dispnet_b200 import loss_functions as LF
    g6 = golden('g6_l1')
    gt = I.sparse_gt(3, 64, 96, seed=110, dataset=ds).to(DEV)
    pred = I.depth_map(3, 64, 96, seed=111, lo=0.0005, hi=95.0 if ds == 'kitti' else 12.0).unsqueeze(1).to(DEV).requires_grad_(True)
    l = LF.l1_loss(gt, [pred], ds)
    assert float(l) == pytest.approx(float(g6[ds]['loss']), rel=1e-5)
    l.backward()
    assert rel(pred.grad, g6[ds]['grad']) < 1e-5
    gt2 = gt.clone()
    gt2[1] = 0           # a sample without valid pixels: mean() of an empty selection is NaN in the reference
    assert math.isnan(float(LF.l1_loss(gt2, [pred.detach()], ds))) and math.isnan(float(g6[ds + '_empty']))


def test_golden_g7_compute_errors_counters_bit_exact(golden):
    from supervised_dispnet_b200 import loss_functions as LF
    g7 = golden('g7_errors')
    gt = I.sparse_gt(3, 128, 416, seed=120, dataset='kitti', density=0.2).to(DEV)
    pred = I.depth_map(3, 128, 416, seed=121, lo=0.0005, hi=95.0).to(DEV)
    for a, b in zip(LF.compute_errors(gt, pred, 'kitti', True), g7['kitti_crop']):
        assert a == pytest.approx(b, rel=1e-5)
    for a, b in zip(LF.compute_errors(gt, pred, 'kitti', True, True), g7['kitti_crop_unsup']):
        assert a == pytest.approx(b, rel=1e-5)
    cnt, _ = LF.error_counters(gt, pred, 'kitti', True)
    assert cnt.tolist() == g7['kitti_crop_counters']          # integer counters: bit-exact
    gtn = I.sparse_gt(2, 64, 96, seed=122, dataset='nyu', density=0.9).to(DEV)
    predn = I.depth_map(2, 64, 96, seed=123, lo=0.0005, hi=12.0).to(DEV)
    for a, b in zip(LF.compute_errors(gtn, predn, 'nyu', False), g7['nyu']):
        assert a == pytest.approx(b, rel=1e-5)
    with pytest.raises(UnboundLocalError):                    # the reference's latent bug is an error here too
        LF.compute_errors(gt, pred, 'kitti', False)


def test_golden_g9_layers_terms(golden):
    """SSIM / edge-aware smoothness / compute_depth_errors (reference layers.py) against reference-generated values."""
    from supervised_dispnet_b200 import layers as LY
    g9 = golden('g9_layers')
    x = (I.images(2, 32, 48, seed=130) * 0.5 + 0.5).to(DEV)
    y = (I.images(2, 32, 48, seed=131) * 0.5 + 0.5).to(DEV)
    assert rel(LY.SSIM()(x, y), g9['ssim']) < 1e-5
    disp = I.depth_map(2, 32, 48, seed=132).unsqueeze(1).to(DEV)
    assert float(LY.get_smooth_loss(disp, x)) == pytest.approx(float(g9['edge_smooth']), rel=1e-5)
    a = I.depth_map(1, 8, 200, seed=133).flatten().to(DEV)
    b = I.depth_map(1, 8, 200, seed=134).flatten().to(DEV)
    for u, v in zip(LY.compute_depth_errors(a, b), g9['depth_errors']):
        assert float(u) == pytest.approx(v, rel=1e-5)


def test_golden_g1_dispnets_eval_config1(golden):
    """BASELINE configs[0]: DispNetS forward on 1x3x128x416 (eval), disparity vs the reference's own output."""
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    m = S.models.DispNetS()
    m.load_state_dict(ON.init_state_dict('DispNetS', 0))
    x = I.images(1, 128, 416, seed=1).to(DEV)
    for prec, tol in (('fp32', 1e-5), ('tc32', 5e-5), ('mixed', 1e-3)):
        m.precision = prec
        m.to(DEV).eval()
        with torch.no_grad():
            d = m(x)
        assert d.shape == (1, 1, 128, 416)
        assert rel(d, golden('g1_dispnets_eval')) < tol, prec


def test_golden_g2_vgg_eval(golden):
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    m = S.models.Disp_vgg_BN()
    m.load_state_dict(ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True), strict=False)
    x = I.images(1, 128, 416, seed=4).to(DEV)
    for prec, tol in (('fp32', 1e-5), ('tc32', 5e-5), ('mixed', 1e-3)):
        m.precision = prec
        m.to(DEV).eval()
        with torch.no_grad():
            d = m(x)
        assert rel(d, golden('g2_vgg_eval')) < tol, prec


@pytest.mark.parametrize('prec,tol', [('fp32', 1e-5), ('tc32', 1e-4)])
def test_golden_g2_vgg_train_fp32(golden, prec, tol):
    """Training-mode forward (4 disparities), BatchNorm running statistics after one step, and parameter gradients
    against the reference-generated fixture (precisions 'fp32' = CUDA cores and 'tc32' = tcgen05 split-bf16)."""
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    g = golden('g2_vgg_train')
    m = S.models.Disp_vgg_BN()
    m.load_state_dict(ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True), strict=False)
    m.precision = prec
    m.to(DEV).train()
    outs = m(I.images(2, 64, 96, seed=3).to(DEV))
    for o, r in zip(outs, g['outs']):
        assert o.shape == r.shape and rel(o, r) < tol
    bufs = dict(m.named_buffers())
    for k, r in g['running'].items():
        if r.dtype.is_floating_point:
            assert rel(bufs[k], r) < tol, k
        else:
            assert int(bufs[k]) == int(r), k
    sum((o * I.probe_like(o, 20 + i).to(DEV)).sum() for i, o in enumerate(outs)).backward()
    named = dict(m.named_parameters())
    for k in g['no_grad_keys']:
        assert named[k].grad is None, k
    num = den = 0.0
    for k, r in g['grads'].items():
        a = I.subsample(named[k].grad.cpu())
        num += float((a.double() - r.double()).norm() ** 2)
        den += float(r.double().norm() ** 2)
    assert math.sqrt(num / den) < (3e-3 if prec == 'fp32' else 1e-2)


# ------------------------------------------------------------------------------------------------------------
# (b) oracle on the same seeded inputs
# ------------------------------------------------------------------------------------------------------------
LOSS_TOL = dict(gx=1e-3, gy=1e-3, loss=1e-5, grad=1e-4, gdepth=1e-3, gpose=1e-3, gmask=1e-3, gimg=1e-3, fwd=1e-4, floats=1e-5, ramp=1e-6, quad=1e-5,
                frac_bad=2e-3)


@pytest.mark.parametrize('name', [n for n, _ in __import__('_parity').LOSS_CASES] if torch.cuda.is_available() else [])
def test_losses_vs_oracle(name):
    fn = dict(P().LOSS_CASES)[name]
    r = fn()
    for k, v in r.items():
        if k in ('counters_equal', 'nan_both'):
            assert v is True, (name, r)
        elif k == 'n_valid':
            assert v > 0
        elif name == 'warp_identity' and k == 'gdepth':
            continue        # analytically zero gradient (depth cancels for t = 0): both sides are round-off noise
        else:
            assert v <= LOSS_TOL[k], (name, k, v)


def _conv_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('gpu_check', os.path.join(os.path.dirname(__file__), '..', 'tools', 'gpu_check.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CONV_CASES


@pytest.mark.parametrize('idx', range(18))
def test_conv_layers_fp32_cuda_core(idx, monkeypatch):
    import _harness as Hn
    monkeypatch.setenv('DISPNET_B200_BACKEND', 'generic')
    name, cfg, shape = _conv_cases()[idx]
    r = Hn.conv_case(cfg, shape, 'fp32')
    for k in ('fwd', 'dx', 'dw', 'db'):
        if k in r:
            assert r[k] < 2e-5, (name, r)


@pytest.mark.parametrize('idx', range(18))
def test_conv_layers_tcgen05(idx):
    """fp16 operands / fp32 TMEM accumulation vs torch fp32 (TF32 off).  Gradients of layers with a fused ReLU /
    LeakyReLU differ where an fp16-rounded pre-activation changes sign (~sqrt(3e-4) of the L2 norm), hence 5e-2 there."""
    import _harness as Hn
    name, cfg, shape = _conv_cases()[idx]
    r = Hn.conv_case(cfg, shape, 'mixed')
    assert r['fwd'] < 1e-3, (name, r)
    gtol = 5e-2 if cfg.get('act', 0) else 6e-3
    for k in ('dx', 'dw', 'db'):
        if k in r:
            assert r[k] < gtol, (name, k, r)
    if cfg.get('stride', 1) == 1 or cfg.get('transposed'):
        assert all(b == 1 for b in r['backends']['fwd']), (name, r['backends'])     # really ran on the tensor cores


@pytest.mark.parametrize('idx', range(18))
def test_conv_layers_tc32(idx):
    """precision 'tc32' (fp32 storage, three split-bf16 tcgen05 terms per product, fp32 TMEM accumulation) vs torch fp32
    (TF32 off): fp32-class results ON the tensor cores.  Gradients of layers with a fused ReLU / LeakyReLU differ where a
    pre-activation within ~1e-5 of zero changes sign (~sqrt(1e-5) of the L2 norm), hence 1e-2 there."""
    import _harness as Hn
    name, cfg, shape = _conv_cases()[idx]
    r = Hn.conv_case(cfg, shape, 'tc32')
    assert r['fwd'] < 3e-5, (name, r)
    gtol = 1e-2 if cfg.get('act', 0) else 3e-5
    for k in ('dx', 'dw', 'db'):
        if k in r:
            assert r[k] < gtol, (name, k, r)
    for kind in ('fwd', 'dgrad', 'wgrad'):
        assert all(b == 1 for b in r['backends'][kind]), (name, r['backends'])      # every GEMM really ran on tcgen05


@pytest.mark.parametrize('name', ['Disp_vgg_BN_train', 'Disp_vgg_BN_eval', 'DispNetS_train', 'PoseExpNet_r2', 'PoseExpNet_r4_exp',
                                  'Disp_res_50_train'])
def test_models_tc32_vs_oracle(name):
    """All four networks in the tensor-core parity mode: outputs within 1e-4 of the fp32 oracle (north star: 1e-3; Disp_res_50
    at 64x96 normalises over 12 samples per channel in layer4 and is allowed 5e-4), BatchNorm running statistics 1e-4,
    gradients within the fp32-vs-fp64 conditioning of the nets themselves."""
    r = dict(P().MODEL_CASES)[name]('tc32')
    assert r['out'] < (5e-4 if 'res_50' in name else 1e-4), r
    if 'running' in r:
        assert r['running'] < (3e-3 if 'res_50' in name else 1e-4), r
    if 'gglobal' in r:
        assert r['gglobal'] < (1e-1 if 'res_50' in name else 1e-2), r


@pytest.mark.parametrize('name', ['Disp_vgg_BN_train', 'Disp_vgg_BN_eval', 'DispNetS_train', 'PoseExpNet_r2', 'PoseExpNet_r4_exp',
                                  'Disp_res_50_train'])
def test_models_fp32_vs_oracle(name, monkeypatch):
    monkeypatch.setenv('DISPNET_B200_BACKEND', 'generic')
    r = dict(P().MODEL_CASES)[name]('fp32')
    assert r['out'] < 1e-5, r
    if 'running' in r:
        assert r['running'] < 1e-4, r
    if 'gglobal' in r:
        # Disp_res_50 at 64x96 normalises over 12 samples per channel in layer4: ill-conditioned in fp32 on both sides
        assert r['gglobal'] < (3e-2 if 'res_50' in name else 3e-3), r


@pytest.mark.parametrize('name', ['Disp_vgg_BN_train', 'Disp_vgg_BN_eval', 'DispNetS_train', 'PoseExpNet_r2', 'PoseExpNet_r4_exp',
                                  'Disp_res_50_train'])
def test_models_tcgen05_vs_oracle(name):
    r = dict(P().MODEL_CASES)[name]('mixed')
    assert r['out'] < (2e-2 if 'res_50' in name else 3e-3), r


@pytest.mark.parametrize('precision', ['tc32', 'mixed'])
def test_disp_res_50_at_configs3_image_size(precision):
    """Disp_res_50 (train mode) at the image size of BASELINE configs[3] (NYU 256x320, batch 4) against the oracle.
      tc32 : measured 1.2e-4 on the outputs (north star 1e-3), BatchNorm running statistics 8e-5; asserted 5e-4.
      mixed: measured 8.5e-3 -- 53 BatchNorm layers deep, the fp16 throughput mode does NOT meet 1e-3 on this network and is
             asserted at the documented 2e-2; the parity mode for Disp_res_50 is tc32."""
    r = P().res50_case(precision, B=4, H=256, W=320)
    if precision == 'tc32':
        assert r['out'] < 5e-4 and r['running'] < 5e-4, r
    else:
        assert r['out'] < 2e-2, r


@pytest.mark.parametrize('precision', ['tc32', 'mixed'])
def test_config2_step_loss_and_disparity_full_size(precision):
    """BASELINE configs[1] at reduced batch: Disp_vgg_BN (train mode) + l1_loss (+0*smooth_loss, train.py:420-522) on
    4x3x128x416 against the oracle.
      tc32  (tensor-core parity mode): all four disparity maps within 1e-4 -- ten times inside the north star's 1e-3 --, loss
             scalar 1e-5, gradients 1e-3 globally and 3e-2 for the worst single parameter (the fp32 reference itself is
             2e-3 away from the fp64 gradient on that parameter, features.features.14.weight).
      mixed (fp16 operands / bf16 gradient activations, the throughput mode): measured 1.4e-4 / 4.6e-4 / 1.2e-3 / 1.5e-3 on the
             four scales -- it does NOT meet 1e-3 on the two coarse maps and is asserted at 3e-3; loss 1e-3; gradients 1e-2
             globally, 0.25 worst parameter."""
    r = P().config2_step_case(precision, B=4, H=128, W=416)
    if precision == 'tc32':
        assert max(r['disp']) < 1e-4, r
        assert r['loss'] < 1e-5 and r['grad_global'] < 1e-3 and r['grad_worst'] < 3e-2, r
    else:
        assert max(r['disp']) < 3e-3 and r['disp'][0] < 1e-3, r
        assert r['loss'] < 1e-3 and r['grad_global'] < 1e-2 and r['grad_worst'] < 0.25, r


def test_no_cpu_fallback():
    import supervised_dispnet_b200 as S
    from supervised_dispnet_b200 import loss_functions as LF
    m = S.models.DispNetS()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 128, 416))
    with pytest.raises(RuntimeError):
        LF.smooth_loss([torch.rand(1, 1, 8, 8)])


def test_cuda_graph_replay_matches_eager(monkeypatch):
    """The captured forward/backward graphs of a plan reproduce the eager launch sequence (same kernels, same buffers):
    five training steps with graphs on and off give the same disparities, loss trajectory and BatchNorm buffers."""
    import supervised_dispnet_b200 as S
    from supervised_dispnet_b200 import loss_functions as LF
    from oracle import nets as ON
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    x = I.images(2, 64, 96, seed=210).to(DEV)
    gt = I.sparse_gt(2, 64, 96, seed=211, dataset='kitti', density=0.3).to(DEV)

    def run(graphs):
        monkeypatch.setenv('DISPNET_B200_GRAPHS', '1' if graphs else '0')
        m = S.models.Disp_vgg_BN()
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
        m.precision = 'mixed'
        m.to(DEV).train()
        opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)
        losses = []
        for _ in range(5):
            d = m(x)
            loss = LF.l1_loss(gt, [1 / t for t in d], 'kitti')
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss))
        plan = m._plan_for([x])
        return losses, d[0].detach().clone(), dict(m.named_buffers())['features.features.1.running_mean'].clone(), plan

    le, de, be, pe = run(False)
    lg, dg, bg, pg = run(True)
    assert pg._fwd_graph is not None and len(pg._bwd_graphs) == 1 and pe._fwd_graph is None
    # weight-gradient partial sums are reduced with fp32 atomics (order varies run to run), so agreement is to round-off
    assert max(abs(a - b) / abs(a) for a, b in zip(le, lg)) < 1e-4
    assert rel(dg, de) < 1e-3 and rel(bg, be) < 1e-4


# ---- disparity-head convolution (warp-MMA kernels) against F.conv2d -----------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('C,H,W,xdt', [(16, 13, 37, 'f16'), (16, 16, 64, 'bf16'), (32, 9, 33, 'f16'), (64, 8, 32, 'f16'),
                                       (128, 5, 20, 'f16'), (16, 24, 70, 'f16')])
def test_head_conv_mma_vs_torch(C, H, W, xdt):
    """nn.Conv2d(C, 1, 3, padding=1) forward / data / weight / bias gradients (reference predict_disp,
    models/Disp_vgg_BN.py:66-70) through dn_head_conv_fwd / dn_head_conv_bwd on ragged tile shapes.  Tolerances: forward
    1e-3 (fp16/bf16 operands, fp32 accumulate); gradients 1e-2 (bf16 operands), relative L2."""
    import ctypes as C_
    import torch
    import torch.nn.functional as F
    from supervised_dispnet_b200 import _lib as L
    torch.manual_seed(C + H)
    dev = torch.device('cuda')
    N = 3
    tdt = torch.float16 if xdt == 'f16' else torch.bfloat16
    ddt = L.DN_F16 if xdt == 'f16' else L.DN_BF16
    x = torch.randn(N, H, W, C, device=dev).to(tdt)
    w = (torch.randn(1, C, 3, 3, device=dev) * 0.2).contiguous()
    b = torch.randn(1, device=dev)
    z = torch.zeros(N, H, W, 1, device=dev)

    def view(t, dt):
        n, h, w_, c = t.shape
        return L.DnView(t.data_ptr(), dt, n, h, w_, c, 0, h * w_ * c, w_ * c, c)
    st = L.stream_ptr()
    vx, vz = view(x, ddt), view(z, L.DN_F32)
    L.call('dn_head_conv_fwd', C_.byref(vx), L.ptr(w), L.ptr(b), C_.byref(vz), st)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=1)
    got = z.permute(0, 3, 1, 2)
    assert float((got - ref).norm() / ref.norm()) < (1e-3 if xdt == 'f16' else 8e-3)

    dz = torch.randn(N, H, W, 1, device=dev)
    ref.backward(dz.permute(0, 3, 1, 2))
    g0 = torch.randn(N, H, W, C, device=dev).bfloat16()
    gx = g0.clone()
    gw, gb = torch.zeros_like(w), torch.zeros(1, device=dev)
    ws = torch.zeros(int(L.lib().dn_reduce_ws_floats(C * 5)), device=dev)
    vdz, vgx = view(dz, L.DN_F32), view(gx, L.DN_BF16)
    L.call('dn_head_conv_bwd', C_.byref(vx), L.ptr(w), C_.byref(vdz), C_.byref(vgx), 1, L.ptr(gw), L.ptr(gb), 1.0, L.ptr(ws), st)
    torch.cuda.synchronize()
    dx = (gx.float() - g0.float()).permute(0, 3, 1, 2)
    assert float((dx - xr.grad).norm() / xr.grad.norm()) < 1.5e-2
    assert float((gw - wr.grad).norm() / wr.grad.norm()) < 1e-2
    assert abs(float(gb) - float(br.grad)) < 1e-3 * (1 + abs(float(br.grad)))
    # overwrite mode (first writer of the gradient slot)
    gx2 = torch.full_like(g0, 7.0)
    vgx2 = view(gx2, L.DN_BF16)
    L.call('dn_head_conv_bwd', C_.byref(vx), L.ptr(w), C_.byref(vdz), C_.byref(vgx2), 0, L.ptr(gw), L.ptr(gb), 1.0, L.ptr(ws), st)
    torch.cuda.synchronize()
    assert float((gx2.float().permute(0, 3, 1, 2) - xr.grad).norm() / xr.grad.norm()) < 1e-2


# ---- BatchNorm(train) + ReLU (+ MaxPool 2x2) kernels against torch, fast paths and generic walkers -----------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('C,H,W,pool,crop', [(64, 8, 12, 0, 0), (64, 8, 12, 1, 0), (96, 6, 10, 1, 0), (200, 5, 7, 0, 0),
                                             (512, 4, 6, 1, 0), (20, 9, 11, 0, 0), (64, 8, 12, 0, 1), (64, 8, 12, 1, 1)])
def test_bn_kernels_vs_torch(C, H, W, pool, crop):
    """dn_bn_train_stats / dn_bn_apply / dn_bn_bwd_reduce / dn_bn_bwd_apply (reference: nn.BatchNorm2d + ReLU + MaxPool2d
    of the vgg16_bn feature stack, models/Disp_vgg_BN.py:137-141) vs torch fp32 on the same fp16-rounded input.  crop=1
    walks a W-cropped view (not pixel-linear), which forces the generic kernels; crop=0 takes the fast paths where the
    channel count allows 16-byte vectors.  Channel counts cover one slab (64), partial slabs (96, 200), many slabs (512)
    and the scalar path (20 channels: no 16-byte vectors).  Tolerances: statistics 1e-5 rel, forward 2e-3 (fp16 output
    rounding), gradients 2e-2 relative L2 (bf16 gradient rounding)."""
    import ctypes as C_
    import torch
    import torch.nn.functional as F
    from supervised_dispnet_b200 import _lib as L
    torch.manual_seed(C * 7 + H)
    dev = torch.device('cuda')
    N, Wb = 3, W + (3 if crop else 0)
    yb = (torch.randn(N, H, Wb, C, device=dev) * 1.5 + 0.3).half()

    def view(t, dt, w=None):
        n, h, wb, c = t.shape
        return L.DnView(t.data_ptr(), dt, n, h, w or wb, c, 0, h * wb * c, wb * c, c)
    y = yb[:, :, :W]
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.2
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    mi, ss = torch.zeros(2 * C, device=dev), torch.zeros(2 * C, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    ws = torch.zeros(int(L.lib().dn_reduce_ws_floats(C)), device=dev)
    st = L.stream_ptr()
    vy = view(yb, L.DN_F16, W)
    for _ in range(2):      # twice: the workspace counters must come back to zero
        L.call('dn_bn_train_stats', C_.byref(vy), L.ptr(gamma), L.ptr(beta), L.ptr(rm), L.ptr(rv), L.ptr(nbt), 0.1, 1e-5, 1,
               L.ptr(sums), L.ptr(mi), L.ptr(ss), L.ptr(ws), st)
    torch.cuda.synchronize()
    assert int(nbt) == 2
    yf = y.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    bn(yf); out_ref = bn(yf)
    mean, var = yf.detach().mean((0, 2, 3)), yf.detach().var((0, 2, 3), unbiased=False)
    assert torch.allclose(mi[:C], mean, rtol=1e-5, atol=1e-5)
    assert torch.allclose(mi[C:], (var + 1e-5).rsqrt(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-4, atol=1e-5) and torch.allclose(rv, bn.running_var, rtol=1e-4, atol=1e-5)
    assert torch.allclose(sums[:C].float(), yf.detach().sum((0, 2, 3)), rtol=1e-4, atol=1e-2)

    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.zeros(N, Ho, Wo, C, device=dev, dtype=torch.float16)
    out2 = torch.zeros(N, Ho, Wo, C, device=dev, dtype=torch.bfloat16)
    vo, vo2 = view(out, L.DN_F16), view(out2, L.DN_BF16)
    L.call('dn_bn_apply', C_.byref(vy), L.ptr(ss), None, L.ACT_RELU, pool, C_.byref(vo), C_.byref(vo2), st)
    ref = F.relu(out_ref)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    got = out.float().permute(0, 3, 1, 2)
    assert float((got - ref).norm() / ref.norm()) < 2e-3
    assert float((out2.float().permute(0, 3, 1, 2) - ref).norm() / ref.norm()) < 8e-3

    g = torch.randn(N, Ho, Wo, C, device=dev).bfloat16()
    ref.backward(g.float().permute(0, 3, 1, 2))
    red = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dyb = torch.zeros(N, H, Wb, C, device=dev, dtype=torch.bfloat16)
    vg, vdy = view(g, L.DN_BF16), view(dyb, L.DN_BF16, W)
    L.call('dn_bn_bwd_reduce', C_.byref(vg), C_.byref(vy), None, L.ptr(mi), L.ptr(gamma), L.ptr(beta), L.ACT_RELU, pool, L.ptr(red),
           L.ptr(ws), st)
    L.call('dn_bn_bwd_apply', C_.byref(vg), C_.byref(vy), None, L.ptr(mi), L.ptr(gamma), L.ptr(beta), L.ACT_RELU, pool, L.ptr(red),
           float(N * H * W), 1.0, L.ptr(dgam), L.ptr(dbet), C_.byref(vdy), None, 0, st)
    torch.cuda.synchronize()
    dy = dyb[:, :, :W].float().permute(0, 3, 1, 2)
    assert float((dy - yf.grad).norm() / yf.grad.norm()) < 2e-2
    assert float((dgam - bn.weight.grad).norm() / bn.weight.grad.norm()) < 5e-3
    assert float((dbet - bn.bias.grad).norm() / bn.bias.grad.norm()) < 5e-3
    assert int(ws[:256].view(torch.int32).abs().sum()) == 0      # arrival counters left at zero


@pytest.mark.gpu
@pytest.mark.parametrize('C,H,W,acc', [(64, 8, 12, 0), (256, 5, 7, 1), (20, 6, 9, 1)])
def test_bn_residual_kernels_vs_torch(C, H, W, acc):
    """out = relu(BN(y) + residual) and its backward (Bottleneck tail, reference models/Disp_res_50.py via torchvision
    resnet50): dy, the residual gradient (overwrite / accumulate) and dgamma / dbeta vs torch.  C = 20 takes the generic
    walkers, the others the fast paths.  Tolerances as in test_bn_kernels_vs_torch."""
    import ctypes as C_
    import torch
    import torch.nn.functional as F
    from supervised_dispnet_b200 import _lib as L
    torch.manual_seed(C + W)
    dev = torch.device('cuda')
    N = 3
    y = (torch.randn(N, H, W, C, device=dev) * 1.5 + 0.3).half()
    r = torch.randn(N, H, W, C, device=dev).half()

    def view(t, dt):
        n, h, w, c = t.shape
        return L.DnView(t.data_ptr(), dt, n, h, w, c, 0, h * w * c, w * c, c)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.2
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    mi, ss = torch.zeros(2 * C, device=dev), torch.zeros(2 * C, device=dev)
    ws = torch.zeros(int(L.lib().dn_reduce_ws_floats(C)), device=dev)
    st = L.stream_ptr()
    vy, vr = view(y, L.DN_F16), view(r, L.DN_F16)
    L.call('dn_bn_train_stats', C_.byref(vy), L.ptr(gamma), L.ptr(beta), L.ptr(rm), L.ptr(rv), None, 0.1, 1e-5, 1, None, L.ptr(mi),
           L.ptr(ss), L.ptr(ws), st)
    out = torch.zeros(N, H, W, C, device=dev, dtype=torch.float16)
    vo = view(out, L.DN_F16)
    L.call('dn_bn_apply', C_.byref(vy), L.ptr(ss), C_.byref(vr), L.ACT_RELU, 0, C_.byref(vo), None, st)
    yf = y.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    rf = r.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    ref = F.relu(bn(yf) + rf)
    assert float((out.float().permute(0, 3, 1, 2) - ref.detach()).norm() / ref.detach().norm()) < 2e-3
    g = torch.randn(N, H, W, C, device=dev).bfloat16()
    ref.backward(g.float().permute(0, 3, 1, 2))
    red = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    dgam, dbet = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dy = torch.zeros(N, H, W, C, device=dev, dtype=torch.bfloat16)
    d0 = torch.randn(N, H, W, C, device=dev).bfloat16()
    dres = d0.clone()
    vg, vdy, vdr = view(g, L.DN_BF16), view(dy, L.DN_BF16), view(dres, L.DN_BF16)
    L.call('dn_bn_bwd_reduce', C_.byref(vg), C_.byref(vy), C_.byref(vr), L.ptr(mi), L.ptr(gamma), L.ptr(beta), L.ACT_RELU, 0, L.ptr(red),
           L.ptr(ws), st)
    L.call('dn_bn_bwd_apply', C_.byref(vg), C_.byref(vy), C_.byref(vr), L.ptr(mi), L.ptr(gamma), L.ptr(beta), L.ACT_RELU, 0, L.ptr(red),
           float(N * H * W), 1.0, L.ptr(dgam), L.ptr(dbet), C_.byref(vdy), C_.byref(vdr), acc, st)
    torch.cuda.synchronize()
    assert float((dy.float().permute(0, 3, 1, 2) - yf.grad).norm() / yf.grad.norm()) < 2e-2
    dr = (dres.float() - (d0.float() if acc else 0)).permute(0, 3, 1, 2)
    assert float((dr - rf.grad).norm() / rf.grad.norm()) < 2e-2
    assert float((dgam - bn.weight.grad).norm() / bn.weight.grad.norm()) < 5e-3
    assert float((dbet - bn.bias.grad).norm() / bn.bias.grad.norm()) < 5e-3


@pytest.mark.parametrize('precision,tol', [('tc32', 1e-4), ('mixed', 5e-3)])
def test_frozen_batchnorm_diff_lr_mode(precision, tol):
    """train.py:405-411 (`--diff_lr`): `disp_net.apply(set_bn_eval)` puts the BatchNorm modules in eval() inside a training
    network -- running statistics normalise (as constants of the graph), the buffers do not move, gamma / beta still train."""
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
    m = S.models.Disp_vgg_BN()
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    m.precision = precision
    m.to(DEV).train()

    def set_bn_eval(mod):
        if mod.__class__.__name__.find('BatchNorm') != -1:
            mod.eval()
    m.apply(set_bn_eval)
    x = I.images(2, 64, 96, seed=11)
    sd_o = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v.clone()) for k, v in sd.items()}
    d_o = ON.disp_vgg_bn(sd_o, x, False)
    (d_o * I.probe_like(d_o, 3)).sum().backward()
    outs = m(x.to(DEV))
    assert len(outs) == 4
    assert rel(outs[0], d_o) < tol
    (outs[0] * I.probe_like(d_o, 3).to(DEV)).sum().backward()
    named = dict(m.named_parameters())
    for k in ('features.features.1.weight', 'features.features.41.bias', 'features.features.0.weight', 'iconv0.0.weight'):
        assert rel(named[k].grad, sd_o[k].grad) < (3e-2 if precision == 'tc32' else 0.3), k     # (ReLU sign flips, as in the train-mode cases)
    bufs = dict(m.named_buffers())
    for k in sd:
        if 'running' in k:
            assert torch.equal(bufs[k].cpu(), sd[k]), k
        if 'num_batches' in k:
            assert int(bufs[k]) == int(sd[k]), k


def test_only_train_dec_detaches_the_encoder():
    """models/Disp_vgg_BN.py:148-153: with `only_train_dec` the encoder outputs are detached -- encoder parameters get no
    gradient, decoder gradients are unchanged."""
    import supervised_dispnet_b200 as S
    from oracle import nets as ON
    sd = ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
    x = I.images(2, 64, 96, seed=12).to(DEV)
    grads = []
    for flag in (False, True):
        m = S.models.Disp_vgg_BN()
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
        m.precision = 'tc32'
        m.only_train_dec = flag
        m.to(DEV).train()
        outs = m(x)
        sum((o * I.probe_like(o, 20 + i).to(DEV)).sum() for i, o in enumerate(outs)).backward()
        grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()})
    full, dec = grads
    for k, g in dec.items():
        if k.startswith('features.'):
            assert g is None, k
        else:
            assert g is not None and rel(g, full[k]) < 1e-6, k


def _g10():
    import test_oracle_golden as TG
    return TG


@pytest.mark.parametrize('name,ds,hi', __import__('test_oracle_golden').G10_CASES)
def test_golden_g10_supervised_losses(golden, name, ds, hi):
    """l2 / berhu / Scale_invariant and the Multiscale_* losses of train.py's `--loss` switch
    (loss_functions.py:77-315) on the masked-reduce kernel family, values and gradients vs the reference's own outputs."""
    from supervised_dispnet_b200 import loss_functions as LF
    g = golden('g10_supervised_losses')[name]
    gt, preds = _g10().g10_inputs(ds, hi, DEV)
    l = _g10().g10_call(LF, name, gt, preds, ds, False)
    l.backward()
    assert abs(float(l) - float(g['loss'])) <= 1e-5 * abs(float(g['loss'])), (float(l), float(g['loss']))
    for p, r in zip(preds, g['grads']):
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            assert rel(p.grad, r) < 1e-4, name
    # deterministic reductions: a second evaluation gives the same bits
    l2 = _g10().g10_call(LF, name, gt, [p.detach() for p in preds], ds, False)
    assert float(l2) == float(l)


def test_berhu_nyu_raises_like_the_reference(golden):
    from supervised_dispnet_b200 import loss_functions as LF
    assert golden('g10_supervised_losses')['berhu_nyu_raises'] == 'UnboundLocalError'
    gt, preds = _g10().g10_inputs('nyu', 12.0, DEV)
    with pytest.raises(UnboundLocalError):
        LF.berhu_loss(gt, preds, 'nyu')


def test_device_input_pipeline_bit_exact_vs_reference_transforms(monkeypatch):
    """uint8 HWC frames -> normalised fp32 NCHW with the per-sample horizontal flip, ground-truth mirror and
    intrinsics fix-up on the device, against the reference's own per-sample chain
    Compose([RandomHorizontalFlip(), ArrayToTensor(), Normalize(.5, .5)]) (custom_transforms.py:14-72, train.py:137-142)."""
    import random
    import numpy as np
    import supervised_dispnet_b200 as S
    from oracle import refshim as R
    root = R.find_root()
    assert root is not None
    ct = R.import_reference(root, with_train=False).custom_transforms
    B, Hh, Ww = 6, 32, 52
    rs = np.random.RandomState(0)
    tgt = rs.randint(0, 256, (B, Hh, Ww, 3)).astype(np.uint8)
    refs = [rs.randint(0, 256, (B, Hh, Ww, 3)).astype(np.uint8) for _ in range(2)]
    gt = rs.rand(B, Hh, Ww).astype(np.float32) * 80
    K = np.tile(np.array([[241.67, 0, 20.4], [0, 246.28, 15.9], [0, 0, 1]], np.float32), (B, 1, 1))
    ref_t = ct.Compose([ct.RandomHorizontalFlip(), ct.ArrayToTensor(), ct.Normalize([0.5] * 3, [0.5] * 3)])
    draws = [0.1, 0.9, 0.3, 0.7, 0.49, 0.51]
    want_img, want_gt, want_K = [], [], []
    for b in range(B):
        monkeypatch.setattr(random, 'random', lambda b=b: draws[b])
        imgs, g, k = ref_t([tgt[b].astype(np.float32)] + [r[b].astype(np.float32) for r in refs], gt[b], K[b])
        want_img.append(imgs)
        want_gt.append(g)
        want_K.append(torch.as_tensor(np.array(k)))
    monkeypatch.undo()
    dt = S.custom_transforms.DeviceTransform(S.custom_transforms.Compose([S.custom_transforms.RandomHorizontalFlip(),
                                                                         S.custom_transforms.ArrayToTensor(),
                                                                         S.custom_transforms.Normalize([0.5] * 3, [0.5] * 3)]), device=DEV)
    outs, g_dev, K_dev = dt([tgt] + refs, gt, K, flips=[int(d < 0.5) for d in draws])
    for j, o in enumerate(outs):
        w = torch.stack([want_img[b][j] for b in range(B)])
        assert o.shape == w.shape and torch.equal(o.cpu(), w), j            # bit-exact
    assert torch.equal(g_dev.cpu(), torch.stack(want_gt))
    assert torch.equal(K_dev.cpu(), torch.stack(want_K))
    # the random draw follows the reference's rule (one random.random() < 0.5 per sample)
    dt2 = S.custom_transforms.DeviceTransform([0.5] * 3, [0.5] * 3, flip=True, device=DEV, rng=random.Random(1))
    r = random.Random(1)
    assert dt2.draw_flips(8) == [int(r.random() < 0.5) for _ in range(8)]
    # the network consumes the transformed batch directly
    m = S.models.Disp_vgg_BN().to(DEV).eval()
    with torch.no_grad():
        assert m(torch.nn.functional.pad(outs[0], (0, 12, 0, 32))).shape == (B, 1, 64, 64)


@pytest.mark.parametrize('name', ['photo_euler_zeros', 'photo_quat_zeros_mask', 'photo_euler_border_mask_r4'])
def test_photometric_pairwise_path_still_matches_oracle(name, monkeypatch):
    """The one-launch-per-(scale, reference) form of the photometric loss (used for pyramid shapes the three-launch path does
    not serve) against the oracle, and bit-level run-to-run determinism of the batched path's scalar."""
    from supervised_dispnet_b200 import loss_functions as LF
    monkeypatch.setattr(LF, 'BATCHED_PHOTO', False)
    r = dict(P().LOSS_CASES)[name]()
    for k, v in r.items():
        assert v <= LOSS_TOL[k], (name, k, v)
    monkeypatch.setattr(LF, 'BATCHED_PHOTO', True)
    tgt, refs, K, Kinv, depth, masks, pose = P()._photo_inputs(2, 2, 64, 96, 0, False)
    args = (tgt.to(DEV), [t.to(DEV) for t in refs], K.to(DEV), Kinv.to(DEV), [d.to(DEV) for d in depth], masks, pose.to(DEV))
    a, b = float(LF.photometric_reconstruction_loss(*args)), float(LF.photometric_reconstruction_loss(*args))
    assert a == b
