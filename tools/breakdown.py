"""Per-kernel-class CUDA-event breakdown of one training step of a model (graphs off):  python tools/breakdown.py dispnets|res50|vgg"""
import os, sys
os.environ['DISPNET_B200_GRAPHS'] = '0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _inputs as I
import supervised_dispnet_b200 as S
from supervised_dispnet_b200 import _lib as L, loss_functions as LF
kind = sys.argv[1] if len(sys.argv) > 1 else 'dispnets'
B = 32
if kind == 'res50':
    H, W = 256, 320
    net = S.models.Disp_res_50('nyu')
else:
    H, W = 128, 416
    net = S.models.DispNetS('kitti') if kind == 'dispnets' else S.models.Disp_vgg_BN('kitti')
net.init_weights(); net.cuda().train()
x = I.images(B, H, W, 1).cuda()
gt = I.sparse_gt(B, H, W, 2, 'kitti').cuda()
opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-4, fused=True)
def step():
    d = net(x); depth = [1 / t for t in d]
    loss = LF.l1_loss(gt, depth, 'kitti')
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
L.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
agg, per = {}, {}
for name, tag, a, b in L.PROFILE:
    key = name if tag is None else '%s[%s,%s]' % (name, tag[0], 'tc' if tag[1] else 'cuda-core')
    t, n = agg.get(key, (0.0, 0)); agg[key] = (t + a.elapsed_time(b), n + 1)
    if tag is not None:
        k = '%s %s %s' % (tag[3], tag[0], 'tc' if tag[1] else 'cc')
        t, f = per.get(k, (0.0, 0.0)); per[k] = (t + a.elapsed_time(b), f + tag[2])
L.PROFILE = None
print('step (events incl. host gaps): %.2f ms; sum of kernels %.2f ms' % (e0.elapsed_time(e1), sum(t for t, n in agg.values())))
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print('  %-40s %4d launches %8.3f ms' % (k, n, t))
print('top layers:')
for k, (t, f) in sorted(per.items(), key=lambda kv: -kv[1][0])[:22]:
    print('  %-40s %8.3f ms %8.2f GFLOP %7.1f TFLOP/s' % (k, t, f / 1e9, f / t / 1e9 if t else 0))
