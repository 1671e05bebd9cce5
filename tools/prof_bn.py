"""Times the HBM-bound BatchNorm / activation kernels on the Disp_vgg_BN shapes of configs[1] (b=32, 128x416) through
the C ABI, in isolation (CUDA events, L2 flushed between launches), and prints achieved GB/s against the algorithmic
bytes.  python tools/prof_bn.py"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from supervised_dispnet_b200 import _lib as L

dev = torch.device('cuda')
B = 32
SHAPES = [(64, 128, 416, 0), (64, 128, 416, 1), (128, 64, 208, 0), (128, 64, 208, 1), (256, 32, 104, 0), (256, 32, 104, 1),
          (512, 16, 52, 0), (512, 16, 52, 1), (512, 8, 26, 0), (512, 8, 26, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def view(t):
    n, h, w, c = t.shape
    dt = {torch.float32: L.DN_F32, torch.float16: L.DN_F16, torch.bfloat16: L.DN_BF16}[t.dtype]
    return L.DnView(t.data_ptr(), dt, n, h, w, c, 0, h * w * c, w * c, c)


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


tot = {}
for (c, h, w, pool) in SHAPES:
    y = torch.randn(B, h, w, c, device=dev).half()
    ho, wo = (h // 2, w // 2) if pool else (h, w)
    out = torch.empty(B, ho, wo, c, device=dev, dtype=torch.float16)
    out2 = torch.empty(B, ho, wo, c, device=dev, dtype=torch.bfloat16)
    gout = torch.randn(B, ho, wo, c, device=dev).bfloat16()
    gy = torch.empty(B, h, w, c, device=dev, dtype=torch.bfloat16)
    gam, bet = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    mi, ss = torch.zeros(2 * c, device=dev), torch.zeros(2 * c, device=dev)
    red = torch.zeros(2 * c, dtype=torch.float64, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    ws = torch.zeros(int(L.lib().dn_reduce_ws_floats(c)), device=dev)
    vy, vo, vo2, vgo, vgy = view(y), view(out), view(out2), view(gout), view(gy)
    s = L.stream_ptr()
    nin, nout = y.numel(), out.numel()
    cases = [
        ('bn_train_stats', lambda: L.call('dn_bn_train_stats', C.byref(vy), L.ptr(gam), L.ptr(bet), L.ptr(rm), L.ptr(rv), L.ptr(nbt), 0.1,
                                          1e-5, 1, None, L.ptr(mi), L.ptr(ss), L.ptr(ws), s), 2 * nin),
        ('bn_apply', lambda: L.call('dn_bn_apply', C.byref(vy), L.ptr(ss), None, L.ACT_RELU, pool, C.byref(vo), C.byref(vo2), s),
         2 * nin + 4 * nout),
        ('bn_bwd_reduce', lambda: L.call('dn_bn_bwd_reduce', C.byref(vgo), C.byref(vy), None, L.ptr(mi), L.ptr(gam), L.ptr(bet), L.ACT_RELU,
                                         pool, L.ptr(red), L.ptr(ws), s), 2 * nin + 2 * nout),
        ('bn_bwd_apply', lambda: L.call('dn_bn_bwd_apply', C.byref(vgo), C.byref(vy), None, L.ptr(mi), L.ptr(gam), L.ptr(bet), L.ACT_RELU,
                                        pool, L.ptr(red), float(B * h * w), 1.0, L.ptr(dg), L.ptr(db), C.byref(vgy), None, 0, s),
         4 * nin + 2 * nout),
    ]
    for name, fn, nbytes in cases:
        fn(); torch.cuda.synchronize()
        ms = timeit(fn)
        tot[name] = tot.get(name, 0) + ms * (2 if (c >= 256 and not pool) else 1)      # layers of that shape in Disp_vgg_BN
        print('%-16s C=%3d %3dx%3d pool=%d  %7.1f us  %6.1f MB  %5.2f TB/s' % (name, c, h, w, pool, ms * 1e3, nbytes / 1e6, nbytes / ms / 1e9))
print({k: round(v, 3) for k, v in tot.items()})
