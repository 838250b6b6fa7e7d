"""Prints the headline numbers of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print('C5', round(d['value'], 2), 'Gbp/s', round(d['ms_per_step'], 3), 'ms digest', d['parity']['digest_ok'], 'k_onesweep frac', round(d['roofline']['frac'], 3),
      'path frac', round(d['path_roofline']['frac'], 3), d['stages_ms'], 'e2e ms', round(d['e2e']['ms_per_step'], 2), 'launches/step', d['gpu_launches'] // d['steps'])
for k, v in d.get('extra', {}).items():
    print(k, round(v['value'], 2), round(v['ms_per_step'], 3), v['digest_ok'], v['stages_ms'], round(v['path_roofline']['frac'], 3), round(v['k_onesweep_frac'] or 0, 3))
