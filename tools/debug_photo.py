import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _inputs as I
from supervised_dispnet_b200 import loss_functions as LF
from oracle import losses as OL
g4 = torch.load(os.path.join(ROOT, 'tests/golden/g4_photometric.pt'), weights_only=False)
DEV = 'cuda'
B, H, W = 2, 64, 96
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())
for R, use_mask in ((2, False), (4, True)):
    tgt = I.images(B, H, W, seed=60); refs = [I.images(B, H, W, seed=61 + r) for r in range(R)]
    K, Kinv = I.intrinsics(B, H / 128.0)
    for rot, pad in (('euler', 'zeros'), ('quat', 'border')):
        mk = lambda dev: ([I.depth_map(B, H >> s, W >> s, seed=70 + s).unsqueeze(1).to(dev).requires_grad_(True) for s in range(4)],
                          I.poses(B, R, seed=80).to(dev).requires_grad_(True),
                          [I.mask_map(B, R, H >> s, W >> s, seed=90 + s).to(dev).requires_grad_(True) for s in range(4)] if use_mask else [None] * 4)
        dp, pp, mp = mk(DEV); do, po, mo = mk('cpu')
        g = g4['R%d_%s_%s' % (R, rot, pad)]
        lp = LF.photometric_reconstruction_loss(tgt.to(DEV), [r.to(DEV) for r in refs], K.to(DEV), Kinv.to(DEV), dp, mp, pp, rot, pad); lp.backward()
        lo = OL.photometric_reconstruction_loss(tgt, refs, K, Kinv, do, mo, po, rot, pad); lo.backward()
        for s in range(4):
            a, b, c = dp[s].grad.cpu(), g['gdepth'][s], do[s].grad
            d = (a - b).abs()
            print(R, rot, pad, 'scale', s, 'prod-vs-golden %.2e oracle-vs-golden %.2e prod-vs-oracle %.2e' % (rel(a, b), rel(c, b), rel(a, c)),
                  'nbad', int((d > 1e-3 * b.abs().max()).sum()), 'max|g|', float(b.abs().max()), 'maxdiff', float(d.max()))
