"""prints the headline numbers and the per-kernel times of bench.py JSON lines"""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith('{'):
            continue
        d = json.loads(l)
        ks = d.pop('kernels', {})
        print(path, d['config']['workload'], 'value', round(d['value'], 1), 'enc', round(d['encode_gbs'], 1), 'dec', round(d['decode_gbs'], 1),
              'ms/step', round(d['ms_per_step'], 2), 'e2e', d.get('e2e') and round(d['e2e']['value'], 1),
              'cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'])
        for k, v in list(ks.items())[:12]:
            print('    %-28s %9.3f ms x%d %s' % (k, v['avg_ms'], v['launches'], ('%.0f GB/s %.3f' % (v['gbs'], v['frac'])) if 'gbs' in v else ''))
