"""Device-resident and train.train() throughput of the other BASELINE configurations (bench.py measures configs[1]):
configs[0] DispNetS + L1, configs[2] Disp_vgg_BN + PoseExpNet photometric, configs[3] Disp_res_50 + L1 (NYU 256x320),
configs[4] DispNetS + PoseExpNet(4, masks) joint.  Synthetic inputs, random-init weights, Adam, CUDA events.
    python tools/bench_configs.py [batch] [config indices, e.g. 2,4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py 32 2,4
(one process per GPU, per-GPU batch fixed = weak scaling, gradients all-reduced by dist.attach; time = max over ranks)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _inputs as I
import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import train as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
WHICH = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 2, 3, 4]
WORLD, RANK, LOCAL = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    from supervised_dispnet_b200 import dist as D
    dist.init_process_group('nccl', device_id=torch.device('cuda', LOCAL))


def build(kind):
    pose = None
    if kind == 'configs[0] DispNetS + L1 (128x416)':
        H, W = 128, 416
        net = S.models.DispNetS('kitti'); net.init_weights()
        batch = (I.images(B, H, W, 1), I.sparse_gt(B, H, W, 2, 'kitti'))
        args = T.default_args(batch_size=B, network='dispnet')
    elif kind == 'configs[2] Disp_vgg_BN + PoseExpNet(2) photometric + smooth (128x416)':
        H, W = 128, 416
        net = S.models.Disp_vgg_BN('kitti'); net.init_weights()
        pose = S.models.PoseExpNet(2, False); pose.init_weights()
        K, Kinv = I.intrinsics(B)
        batch = (I.images(B, H, W, 1), [I.images(B, H, W, 3 + r) for r in range(2)], K, Kinv, None)
        args = T.default_args(batch_size=B, unsupervised=True, smooth_loss_weight=0.1)
    elif kind == 'configs[3] Disp_res_50 + L1 (NYU 256x320)':
        H, W = 256, 320
        net = S.models.Disp_res_50('nyu'); net.init_weights()
        gt = torch.stack([I.sparse_gt(B, H, W, 2, 'nyu', density=0.9), torch.ones(B, H, W)], 1)
        batch = (I.images(B, H, W, 1), gt)
        args = T.default_args(batch_size=B, dataset='nyu', network='disp_res_50')
    else:
        H, W = 128, 416
        net = S.models.DispNetS('kitti'); net.init_weights()
        pose = S.models.PoseExpNet(4, True); pose.init_weights()
        K, Kinv = I.intrinsics(B)
        batch = (I.images(B, H, W, 1), [I.images(B, H, W, 3 + r) for r in range(4)], K, Kinv, None)
        args = T.default_args(batch_size=B, unsupervised=True, smooth_loss_weight=0.1, mask_loss_weight=0.2)
    return net, pose, batch, args


def pin(o):
    if torch.is_tensor(o):
        return o.pin_memory()
    if isinstance(o, (list, tuple)):
        return type(o)(pin(x) for x in o)
    return o


KINDS = ['configs[0] DispNetS + L1 (128x416)', 'configs[2] Disp_vgg_BN + PoseExpNet(2) photometric + smooth (128x416)',
         'configs[3] Disp_res_50 + L1 (NYU 256x320)', 'configs[4] DispNetS + PoseExpNet(4, masks) joint (128x416)']
for kind in [k for k in KINDS if int(k[8]) in WHICH]:
    net, pose, batch, args = build(kind)
    net.cuda()
    params = [p for p in net.parameters() if p.requires_grad]
    if pose is not None:
        pose.cuda(); params += list(pose.parameters())
    if WORLD > 1:
        D.attach(net)
        if pose is not None:
            D.attach(pose)
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    hb = pin(batch)
    T.train(args, [hb] * (6 if WORLD == 1 else 20), net, pose, opt, 6 if WORLD == 1 else 20)
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    K = 15 if WORLD == 1 else 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = T.train(args, [hb] * K, net, pose, opt, K)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device='cuda')
    if WORLD > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    if RANK == 0:
        print('%-78s N=%d b=%d/GPU  %7.2f ms/step  %8.1f images/s  (train.train, pinned host batches; avg loss %.4f)' % (
            kind, WORLD, B, ms, WORLD * B / ms * 1e3, loss), flush=True)
    del net, pose, opt
    torch.cuda.empty_cache()

if WORLD > 1:
    dist.destroy_process_group()
