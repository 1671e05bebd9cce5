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
