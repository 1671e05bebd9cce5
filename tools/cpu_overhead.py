"""Host-side enqueue time of one training step (no sync inside the loop) vs device time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import loss_functions as LF
net = S.models.Disp_vgg_BN('kitti'); net.init_weights(); net = net.cuda().train()
opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-4, fused=True)
x, gt = bench.synth_batch(32, 10); x, gt = x.cuda(), gt.cuda()
def step():
    disp = net(x); depth = [1 / d for d in disp]
    loss = LF.l1_loss(gt, depth, 'kitti') + 0.0 * LF.smooth_loss(depth)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('enqueue %.2f ms/step, total %.2f ms/step' % ((t1 - t0) * 100, (t2 - t0) * 100))

# the same through train.train() with pinned host batches (prefetcher + deferred loss read)
from supervised_dispnet_b200 import train as T
xh, gh = bench.synth_batch(32, 10, pinned=True) if 'pinned' in bench.synth_batch.__code__.co_varnames else (x.cpu().pin_memory(), gt.cpu().pin_memory())
targs = T.default_args(batch_size=32, smooth_loss_weight=0.0)
T.train(targs, [(xh, gh)] * 3, net, None, opt, 3)
torch.cuda.synchronize()
K = 20
t0 = time.perf_counter()
T.train(targs, [(xh, gh)] * K, net, None, opt, K)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('train.train: host returns after %.2f ms/step, total %.2f ms/step' % ((t1 - t0) * 1e3 / K, (t2 - t0) * 1e3 / K))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
T.train(targs, [(xh, gh)] * K, net, None, opt, K)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
