"""Prints the configs[1] full-size parity figures (disparities, loss, gradients vs the fp32 and fp64 oracle) per precision."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import _parity as P  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for prec in sys.argv[2:] or ['tc32', 'mixed']:
    r = P.config2_step_case(prec, B=B, fp64=True)
    print(prec, json.dumps({k: (v if not isinstance(v, float) else float('%.3e' % v)) for k, v in r.items()}), flush=True)
