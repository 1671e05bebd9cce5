"""Loss-path workload at the benchmark's size (b=32, 128x416, R=2 reference frames, 4 scales): photometric_reconstruction_loss
+ smooth_loss + l1_loss forward and backward, a few iterations, with CUDA-event timing per call group.  Run under ncu for the
launch list / full captures of warp_photo_*, smooth_*, l1 (profiles/README.md)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import _inputs as I  # noqa: E402
from supervised_dispnet_b200 import loss_functions as LF  # noqa: E402

B, H, W, R = 32, 128, 416, 2
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = 'cuda'
tgt = I.images(B, H, W, 1).to(dev)
refs = [I.images(B, H, W, 2 + i).to(dev) for i in range(R)]
K, Kinv = I.intrinsics(B)
K, Kinv = K.to(dev), Kinv.to(dev)
depth = [I.depth_map(B, H >> s, W >> s, 10 + s).unsqueeze(1).to(dev).requires_grad_(True) for s in range(4)]
pose = I.poses(B, R, 30).to(dev).requires_grad_(True)
gt = I.sparse_gt(B, H, W, 40, 'kitti').to(dev)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def photo():
    l = LF.photometric_reconstruction_loss(tgt, refs, K, Kinv, depth, [None] * 4, pose, 'euler', 'zeros')
    l.backward()


def smooth():
    LF.smooth_loss(depth).backward()


def l1():
    LF.l1_loss(gt, depth, 'kitti').backward()


hw = sum((H >> s) * (W >> s) for s in range(4))
# algorithmic bytes: fwd reads tgt + R refs + depth per scale, bwd reads them again and writes the depth gradient
photo_bytes = 4 * B * (2 * (3 * (1 + R) + 1) * hw + hw) + 4 * B * 3 * (1 + R) * H * W      # + the pyramid's read of the full-size frames
for name, fn, nbytes in (('photometric fwd+bwd (R=2, 4 scales)', photo, photo_bytes), ('smooth fwd+bwd', smooth, 4 * B * 3 * hw),
                         ('l1 fwd+bwd', l1, 4 * B * 3 * H * W)):
    ms = timed(fn, iters)
    print('%-40s %8.3f ms  algorithmic %7.1f MB  -> %7.1f GB/s' % (name, ms, nbytes / 1e6, nbytes / ms / 1e6), flush=True)
