"""Prints the interesting parts of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
for k in ('value', 'ms_per_step', 'dtype', 'fill_only_gcups', 'roofline_frac_solve_plus_traceback', 'parity_spotcheck',
          'e2e_byte_forms', 'gpu_launches', 'clocks'):
    print(k, json.dumps(d.get(k)))
print('cpu', json.dumps(d.get('cpu_baseline')))
e = d['e2e']
print('e2e', e['value'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e.get('breakdown_ms'))
print('roofline', d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['peak'])
for k, v in (d.get('other_workloads') or {}).items():
    print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ('workload', 'e2e_breakdown_ms', 'e2e_form')})
    print('    ', v.get('e2e_breakdown_ms'))
print('sharded', json.dumps(d.get('sharded_batch')))
