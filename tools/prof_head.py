"""Times dn_head_conv_fwd / dn_head_conv_bwd on the four Disp_vgg_BN head shapes of configs[1] (b=32), L2 flushed."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from supervised_dispnet_b200 import _lib as L
dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def view(t, dt):
    n, h, w, c = t.shape
    return L.DnView(t.data_ptr(), dt, n, h, w, c, 0, h * w * c, w * c, c)


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for (c, h, w) in [(16, 128, 416), (32, 64, 208), (64, 32, 104), (128, 16, 52)]:
    x = torch.randn(32, h, w, c, device=dev).half()
    wt = torch.randn(1, c, 3, 3, device=dev) * 0.1
    b = torch.zeros(1, device=dev)
    z = torch.zeros(32, h, w, 1, device=dev)
    dz = torch.randn(32, h, w, 1, device=dev).bfloat16()
    gx = torch.zeros(32, h, w, c, device=dev, dtype=torch.bfloat16)
    gw, gb = torch.zeros_like(wt), torch.zeros(1, device=dev)
    ws = torch.zeros(int(L.lib().dn_reduce_ws_floats(c * 5)), device=dev)
    vx, vz, vdz, vgx = view(x, L.DN_F16), view(z, L.DN_F32), view(dz, L.DN_BF16), view(gx, L.DN_BF16)
    s = L.stream_ptr()
    f = lambda: L.call('dn_head_conv_fwd', C.byref(vx), L.ptr(wt), L.ptr(b), C.byref(vz), s)
    g = lambda: L.call('dn_head_conv_bwd', C.byref(vx), L.ptr(wt), C.byref(vdz), C.byref(vgx), 1, L.ptr(gw), L.ptr(gb), 1.0, L.ptr(ws), s)
    f(); g(); torch.cuda.synchronize()
    tf, tb = timeit(f), timeit(g)
    nb = x.numel() * 2
    print('C=%3d %3dx%3d  fwd %6.1f us (%4.2f TB/s)   bwd %6.1f us (%4.2f TB/s incl. gx RMW)' % (c, h, w, tf * 1e3, (nb + z.numel() * 4) / tf / 1e9, tb * 1e3, 3 * nb / tb / 1e9))
