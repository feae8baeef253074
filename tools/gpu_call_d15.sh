#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --points 2048 --k 40 --batch 32 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/d16_cfg2_n1.json 2> gpurun_out/d16_cfg2_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d16_cfg2_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['config']['workload'])
print('roofline',d['roofline']['frac'], d['roofline']['us_per_step'])
print('edgeconv',d['roofline_edgeconv'])
for k,v in list(d['kernel_ms_per_step'].items())[:22]: print('  ',k[:100],v)
PY
