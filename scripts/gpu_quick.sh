#!/bin/bash
# one GPU call: the parity suite, then a short bench line (kernel times); outputs under gpurun_out/<tag>_*
tag=${1:-q}; shift
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo rc=$? >> gpurun_out/${tag}_tests.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-index "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_tests.log
python - <<P
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print('step',round(d['ms_per_step'],2),'enc',round(d['encode_gbs']),'dec',round(d['decode_gbs']))
for k,v in d['kernels'].items():
    if v['avg_ms']>0.05: print(' ',k, round(v['avg_ms'],3), v.get('parts_avg_ms',''))
P
