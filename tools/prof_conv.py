"""Runs a few full-size single-layer cases (forward + backward) for ncu captures / event timing.
    python tools/prof_conv.py [case ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _harness as Hn
CASES = {
    'iconv0': (dict(cin=17, cout=16, k=3, act=2), (32, 17, 128, 416)),
    'feat3': (dict(cin=64, cout=64, k=3), (32, 64, 128, 416)),
    'feat0': (dict(cin=3, cout=64, k=3), (32, 3, 128, 416)),
    'feat7': (dict(cin=64, cout=128, k=3), (32, 64, 64, 208)),
    'iconv2': (dict(cin=193, cout=64, k=3, act=2), (32, 193, 32, 104)),
    'iconv3': (dict(cin=385, cout=128, k=3, act=2), (32, 385, 16, 52)),
    'feat10': (dict(cin=128, cout=128, k=3), (32, 128, 64, 208)),
    'feat17': (dict(cin=256, cout=256, k=3), (32, 256, 32, 104)),
    'feat27': (dict(cin=512, cout=512, k=3), (32, 512, 16, 52)),
    'upconv0': (dict(cin=32, cout=16, k=4, stride=2, pad=1, transposed=True, act=2), (32, 32, 64, 208)),
    'upconv1': (dict(cin=64, cout=32, k=4, stride=2, pad=1, transposed=True, act=2), (32, 64, 32, 104)),
    'iconv1': (dict(cin=97, cout=32, k=3, act=2), (32, 97, 64, 208)),
}
names = sys.argv[1:] or list(CASES)
prec = os.environ.get('DISPNET_B200_PRECISION', 'mixed')
for n in names:
    cfg, shape = CASES[n]
    torch.manual_seed(0)
    m = Hn.OneConv(precision=prec, **cfg).cuda().train()
    x = torch.randn(shape, device='cuda')
    for it in range(3):
        out = m(x)
        out.sum().backward()
    torch.cuda.synchronize()
    from supervised_dispnet_b200 import _lib as L
    L.PROFILE = []
    out = m(x); out.sum().backward(); torch.cuda.synchronize()
    for name, tag, a, b in L.PROFILE:
        if tag:
            print('%-8s %-6s %s %8.3f ms %8.1f TFLOP/s' % (n, tag[0], 'tc' if tag[1] else 'cc', a.elapsed_time(b), tag[2] / a.elapsed_time(b) / 1e9))
    L.PROFILE = None

# optional per-role cycle counters of the tcgen05 kernels: DN_TC_DEBUG=1 (forward + backward of one step, summed over CTAs)
if os.environ.get('DN_TC_DEBUG'):
    import ctypes
    for n in names:
        cfg, shape = CASES[n]
        torch.manual_seed(0)
        m = Hn.OneConv(precision=prec, **cfg).cuda().train()
        x = torch.randn(shape, device='cuda')
        for _ in range(2):
            m(x).sum().backward()
        cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
        L.lib().dn_tc_set_debug(ctypes.c_void_p(cnt.data_ptr()))
        with torch.no_grad():
            m(x)
        torch.cuda.synchronize()
        c = cnt.cpu().tolist()
        nct = 148.0
        print('%-8s fwd   cycles/CTA: producer wait_empty %.0f of %.0f | mma wait_full %.0f wait_tmem_empty %.0f of %.0f | epi wait_full %.0f work %.0f'
              % (n, c[0] / nct, c[1] / nct, c[2] / nct, c[3] / nct, c[4] / nct, c[5] / nct, c[6] / nct))
        cnt.zero_()
        m(x).sum().backward()
        torch.cuda.synchronize()
        L.lib().dn_tc_set_debug(None)
        c = cnt.cpu().tolist()
        print('%-8s   epilogue detail: first TMEM load ready after %.0f, fence+arrive %.0f' % (n, c[7] / nct, c[15] / nct))
        print('%-8s f+dgr cycles/CTA: producer wait_empty %.0f of %.0f | mma wait_full %.0f wait_tmem_empty %.0f of %.0f | epi wait_full %.0f work %.0f'
              % (n, c[0] / nct, c[1] / nct, c[2] / nct, c[3] / nct, c[4] / nct, c[5] / nct, c[6] / nct))
        print('%-8s wgrad cycles/CTA: producer wait_empty %.0f of %.0f | mma wait_full %.0f wait_tmem_empty %.0f of %.0f | epi wait_full %.0f work %.0f'
              % (n, c[8] / nct, c[9] / nct, c[10] / nct, c[11] / nct, c[12] / nct, c[13] / nct, c[14] / nct))
