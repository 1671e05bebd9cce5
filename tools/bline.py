"""Prints the interesting fields of the last JSON line of a bench log: python tools/bline.py gpurun_out/b.log"""
import json
import sys
j = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value %.1f  ms/step %.3f  e2e %.1f  frac %.3f' % (j['value'], j['ms_per_step'], j.get('e2e', {}).get('value', 0), j.get('roofline', {}).get('frac', 0)))
print(j.get('kernel_breakdown_ms_per_step'))
print("clocks", j.get('clocks'), 'launches', j.get('gpu_launches'))
