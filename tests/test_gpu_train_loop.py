"""GPU: the reference's training loop (mirror of train.train, reference train.py:394-539) runs every BASELINE configuration
end to end on the product path at reduced batch: losses are finite, parameters of every trained network move, and the
first-step loss equals the oracle's on the same weights and batch (1e-3)."""
import math

import pytest
import torch

import _inputs as I

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _loader(batch, n):
    return [batch] * n


def _first_loss_oracle(kind, sd_d, sd_p, batch, args):
    from oracle import nets as ON, losses as OL
    if kind == 'vgg_l1':
        x, gt = batch
        d = ON.disp_vgg_bn({k: v.clone() for k, v in sd_d.items()}, x, True)
        return float(OL.l1_loss(gt, [1 / t for t in d], 'kitti'))
    if kind == 'res50_nyu':
        x, gt = batch
        d = ON.disp_res_50({k: v.clone() for k, v in sd_d.items()}, x, True, 'nyu')
        return float(OL.l1_loss(gt[:, 0], [1 / t for t in d], 'nyu'))
    x, refs, K, Kinv, _ = batch
    fwd = ON.disp_vgg_bn if kind == 'vgg_photo' else ON.dispnets
    d = fwd({k: v.clone() for k, v in sd_d.items()}, x, True)
    depth = [1 / t for t in d]
    exp = kind == 'dispnets_joint'
    masks, pose = ON.poseexpnet({k: v.clone() for k, v in sd_p.items()}, x, refs, True, exp)
    l1 = OL.photometric_reconstruction_loss(x, refs, K, Kinv, depth, masks, pose, 'euler', 'zeros')
    l2 = OL.explainability_loss(masks) if args.mask_loss_weight > 0 else 0
    l3 = OL.smooth_loss(depth)
    return float(args.photo_loss_weight * l1 + args.mask_loss_weight * l2 + args.smooth_loss_weight * l3)


@pytest.mark.parametrize('kind', ['vgg_l1', 'vgg_photo', 'res50_nyu', 'dispnets_joint'])
def test_train_loop_configs(kind):
    import supervised_dispnet_b200 as S
    from supervised_dispnet_b200 import train as T
    from oracle import nets as ON
    B = 2
    pose_net, sd_p = None, None
    if kind == 'vgg_l1':           # BASELINE configs[1]
        H, W = 128, 416
        net, sd_d = S.models.Disp_vgg_BN('kitti'), ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
        batch = (I.images(B, H, W, 1), I.sparse_gt(B, H, W, 2, 'kitti'))
        args = T.default_args(batch_size=B)
    elif kind == 'vgg_photo':      # configs[2]: Disp_vgg_BN + PoseExpNet(R=2) photometric + smooth
        H, W = 128, 416
        net, sd_d = S.models.Disp_vgg_BN('kitti'), ON.init_state_dict('Disp_vgg_BN', 0, skip_dead=True)
        pose_net, sd_p = S.models.PoseExpNet(2, False), ON.init_state_dict('PoseExpNet', 1, nb_ref_imgs=2, output_exp=False)
        K, Kinv = I.intrinsics(B)
        batch = (I.images(B, H, W, 1), [I.images(B, H, W, 3 + r) for r in range(2)], K, Kinv, None)
        args = T.default_args(batch_size=B, unsupervised=True, smooth_loss_weight=0.1)
    elif kind == 'res50_nyu':      # configs[3]
        H, W = 256, 320
        net, sd_d = S.models.Disp_res_50('nyu'), ON.init_state_dict('Disp_res_50', 0)
        gt = torch.stack([I.sparse_gt(B, H, W, 2, 'nyu', density=0.9), torch.ones(B, H, W)], 1)
        batch = (I.images(B, H, W, 1), gt)
        args = T.default_args(batch_size=B, dataset='nyu')
    else:                          # configs[4]: DispNetS + PoseExpNet(R=4, masks) joint
        H, W = 128, 416
        net, sd_d = S.models.DispNetS('kitti'), ON.init_state_dict('DispNetS', 0)
        pose_net, sd_p = S.models.PoseExpNet(4, True), ON.init_state_dict('PoseExpNet', 1, nb_ref_imgs=4, output_exp=True)
        K, Kinv = I.intrinsics(B)
        batch = (I.images(B, H, W, 1), [I.images(B, H, W, 3 + r) for r in range(4)], K, Kinv, None)
        args = T.default_args(batch_size=B, unsupervised=True, smooth_loss_weight=0.1, mask_loss_weight=0.2)
    net.load_state_dict({k: v.clone() for k, v in sd_d.items()}, strict=False)
    net.to(DEV)
    params = [p for p in net.parameters() if p.requires_grad]
    if pose_net is not None:
        pose_net.load_state_dict({k: v.clone() for k, v in sd_p.items()})
        pose_net.to(DEV)
        params += list(pose_net.parameters())          # joint optimisation as in the reference's commented block train.py:291-294
    before = [p.detach().clone() for p in params]
    opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.999))
    ref = _first_loss_oracle(kind, sd_d, sd_p, batch, args)
    first = T.train(args, _loader(batch, 1), net, pose_net, opt, 1)
    assert math.isfinite(first) and abs(first - ref) < 2e-3 * abs(ref), (first, ref)
    avg = T.train(args, _loader(batch, 4), net, pose_net, opt, 4)       # crosses the CUDA-graph capture point
    assert math.isfinite(avg)
    moved = sum(int((a - b.detach()).abs().max() > 0) for a, b in zip(before, params))
    dead = 2 if kind == 'res50_nyu' else 0                                # Disp_res_50.bn1 never receives a gradient
    assert moved >= len(params) - dead - 13, (moved, len(params))         # (biases in front of BatchNorm get exact zeros)
