"""Per-call device timing of one forward + backward of a whole network (any of the four), serialised (no graphs / side streams):
    python tools/prof_model.py Disp_res_50 16 256 320     -> time per C-ABI entry point and the slowest GEMM launches"""
import collections
import os
import sys

os.environ.setdefault('DISPNET_B200_GRAPHS', '0')
os.environ.setdefault('DISPNET_B200_SIDE_STREAM', '0')
os.environ.setdefault('DISPNET_B200_PHASE_STREAMS', '0')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import supervised_dispnet_b200 as S  # noqa: E402
from supervised_dispnet_b200 import _lib as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'Disp_res_50'
B, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (16, 256, 320)
net = {'Disp_res_50': lambda: S.models.Disp_res_50('nyu'), 'Disp_vgg_BN': lambda: S.models.Disp_vgg_BN('kitti'),
       'DispNetS': lambda: S.models.DispNetS('kitti')}[name]()
net.init_weights()
net = net.cuda().train()
x = torch.randn(B, 3, H, W, device='cuda')


def step():
    outs = net(x)
    sum(o.sum() for o in outs).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
L.PROFILE = []
step()
torch.cuda.synchronize()
per = collections.defaultdict(float)
rows = []
for nm, tag, a, b in L.PROFILE:
    ms = a.elapsed_time(b)
    per[nm + ('[%s]' % tag[0] if tag else '')] += ms
    if tag:
        rows.append((ms, tag[3], tag[0], tag[2] / ms / 1e9 if ms > 0 else 0.0))
L.PROFILE = None
tot = sum(per.values())
print('%s b=%d %dx%d: %.3f ms of device time in %d calls' % (name, B, H, W, tot, len(rows)))
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:16]:
    print('  %-34s %8.3f ms  %5.1f %%' % (k, v, 100 * v / tot))
print('slowest GEMM launches:')
for ms, lname, kind, tf in sorted(rows, reverse=True)[:24]:
    print('  %-34s %-6s %7.3f ms %8.1f TFLOP/s' % (lname, kind, ms, tf))
