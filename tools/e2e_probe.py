"""Repeated timing of the public train.train() loop (pinned host batches) vs the device-resident loop, to see jitter."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import loss_functions as LF, train as T
net = S.models.Disp_vgg_BN('kitti'); net.init_weights(); net = net.cuda().train()
opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-4, fused=True)
xh, gh = bench.synth_batch(32, 10, pinned=True)
x, gt = xh.cuda(), gh.cuda()
targs = T.default_args(batch_size=32, smooth_loss_weight=0.0)
def step():
    disp = net(x); depth = [1 / d for d in disp]
    loss = LF.l1_loss(gt, depth, 'kitti') + 0.0 * LF.smooth_loss(depth)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(6): step()
T.train(targs, [(xh, gh)] * 4, net, None, opt, 4)
torch.cuda.synchronize()
K = 20
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): step()
    e1.record(); torch.cuda.synchronize()
    d = e0.elapsed_time(e1) / K
    e0.record(); t0 = time.perf_counter()
    T.train(targs, [(xh, gh)] * K, net, None, opt, K)
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    print('rep %d: device loop %.3f ms/step | train.train %.3f ms/step (host loop returned after %.3f)' % (rep, d, e0.elapsed_time(e1) / K, (t1 - t0) * 1e3 / K), flush=True)
# per-step wall time inside train.train: where do the slow steps sit?
import types
stamps = []
class L(list):
    def __iter__(self):
        for b in list.__iter__(self):
            stamps.append(time.perf_counter()); yield b
T.train(targs, L([(xh, gh)] * 40), net, None, opt, 40); torch.cuda.synchronize()
d = [(b - a) * 1e3 for a, b in zip(stamps, stamps[1:])]
print('per-iteration host ms:', ' '.join('%.1f' % v for v in d))
