"""GPU bring-up battery: prints relative errors of the product path against torch (layers) and the oracle (models,
losses) without stopping at the first failure.  Each section runs in its own process under `timeout`
(tools/gpu_check.sh) so that a faulting or hanging kernel cannot take the rest down.

    python tools/gpu_check.py <section> [--precision fp32|fp16|...]
"""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

RESULTS = []


def report(name, **kw):
    RESULTS.append(dict(name=name, **kw))
    print('%-46s %s' % (name, ' '.join('%s=%s' % (k, ('%.3e' % v) if isinstance(v, float) else v) for k, v in kw.items())),
          flush=True)


def guarded(name, fn):
    try:
        t0 = time.time()
        out = fn()
        torch.cuda.synchronize()
        if isinstance(out, dict):
            report(name, **out, sec=round(time.time() - t0, 2))
        else:
            report(name, ok=True, sec=round(time.time() - t0, 2))
    except Exception as e:  # noqa: BLE001
        report(name, error=repr(e)[:300])
        traceback.print_exc()


CONV_CASES = [
    ('c3x3_s1_64_64', dict(cin=64, cout=64, k=3), (2, 64, 16, 24)),
    ('c3x3_s1_3_64', dict(cin=3, cout=64, k=3), (2, 3, 32, 32)),
    ('c3x3_s1_193_64_lrelu', dict(cin=193, cout=64, k=3, act=2), (2, 193, 8, 24)),
    ('c3x3_s1_128_256', dict(cin=128, cout=256, k=3, act=1), (4, 128, 8, 26)),
    ('c3x3_s1_512_512', dict(cin=512, cout=512, k=3), (8, 512, 4, 13)),
    ('c3x3_s1_17_16', dict(cin=17, cout=16, k=3, act=2), (2, 17, 32, 32)),
    ('c3x3_s1_16_1_head', dict(cin=16, cout=1, k=3), (2, 16, 16, 16)),
    ('c7x7_s2_3_32', dict(cin=3, cout=32, k=7, stride=2, act=1), (2, 3, 32, 48)),
    ('c5x5_s2_32_64', dict(cin=32, cout=64, k=5, stride=2, act=1), (2, 32, 16, 24)),
    ('c3x3_s2_64_128', dict(cin=64, cout=128, k=3, stride=2, act=1), (2, 64, 15, 13)),
    ('c1x1_s1_64_256_nobias', dict(cin=64, cout=256, k=1, pad=0, bias=False), (2, 64, 8, 8)),
    ('c1x1_s2_256_512_nobias', dict(cin=256, cout=512, k=1, pad=0, stride=2, bias=False), (2, 256, 8, 8)),
    ('ct4_s2_512_256', dict(cin=512, cout=256, k=4, stride=2, pad=1, transposed=True, act=2), (2, 512, 4, 13)),
    ('ct4_s2_32_16', dict(cin=32, cout=16, k=4, stride=2, pad=1, transposed=True, act=2), (2, 32, 16, 24)),
    ('ct3_s2_op1_512_512_crop', dict(cin=512, cout=512, k=3, stride=2, pad=1, transposed=True, out_pad=1, act=1,
                                     crop=(2, 7)), (2, 512, 1, 4)),
    ('ct3_s2_op1_64_32', dict(cin=64, cout=32, k=3, stride=2, pad=1, transposed=True, out_pad=1, act=1), (2, 64, 8, 12)),
    # thin output gradients: the weight gradient stacks three column-shifted dy copies in the MMA M dimension (dn_tc.cu, mstack)
    ('c3x3_s1_128_4_mstack', dict(cin=128, cout=4, k=3), (4, 128, 8, 12)),
    ('c3x3_s1_97_32_mstack', dict(cin=97, cout=32, k=3, act=2), (2, 97, 16, 24)),
]


def section_conv(precision):
    import _harness as Hn
    for name, cfg, shape in CONV_CASES:
        guarded('conv/%s/%s' % (precision, name), lambda cfg=cfg, shape=shape: Hn.conv_case(cfg, shape, precision))


def section_losses(_):
    import _parity as P
    for name, fn in P.LOSS_CASES:
        guarded('loss/' + name, fn)


def section_models(precision):
    import _parity as P
    for name, fn in P.MODEL_CASES:
        guarded('model/%s/%s' % (precision, name), lambda fn=fn: fn(precision))


def section_full(precision):
    import _parity as P
    guarded('full/%s/Disp_vgg_BN_b4_128x416' % precision, lambda: P.vgg_case(precision, B=4, H=128, W=416))
    guarded('full/%s/Disp_res_50_b2_128x160' % precision, lambda: P.res50_case(precision, B=2, H=128, W=160))
    guarded('full/%s/DispNetS_b4_128x416' % precision, lambda: P.dispnets_case(precision, B=4, H=128, W=416))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('section')
    ap.add_argument('--precision', default='fp32')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    print('== section %s precision %s | torch %s | %s' % (a.section, a.precision, torch.__version__,
                                                           torch.cuda.get_device_name(0)), flush=True)
    from supervised_dispnet_b200 import _lib as L
    print('tc_available', L.lib().dn_tc_available(), 'backend env', os.environ.get('DISPNET_B200_BACKEND', 'auto'), flush=True)
    {'conv': section_conv, 'losses': section_losses, 'models': section_models, 'full': section_full}[a.section](a.precision)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, 'w') as f:
            json.dump(RESULTS, f, indent=1)


if __name__ == '__main__':
    main()
