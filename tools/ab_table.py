"""Table of tools/ab.sh output: per (case, kind) milliseconds of the two builds."""
import collections, re, sys
rows = collections.OrderedDict(); lib = None
for l in open(sys.argv[1]):
    if l.startswith('=='):
        lib = l.split()[1]; continue
    m = re.match(r'(\S+)\s+(fwd|wgrad|dgrad)\s+(tc|cc)\s+([\d.]+) ms', l)
    if m:
        rows.setdefault((m.group(1), m.group(2)), {}).setdefault(lib, []).append(float(m.group(4)))
tb = tn = 0
for k, v in rows.items():
    b = sum(v.get('libdispnet_b200_base.so', [0])); n = sum(v.get('libdispnet_b200.so', [0])); tb += b; tn += n
    print('%-8s %-6s base %.3f new %.3f  %+.0f%%' % (k[0], k[1], b, n, 100 * (n - b) / b if b else 0))
print('total base %.3f new %.3f' % (tb, tn))
