#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/d16_pytest.log; tail -6 gpurun_out/d16_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/d16_bench.json 2> gpurun_out/d16_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d16_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',d['roofline']['frac'], d['roofline']['frac_of_3xtf32_roofline'], d['roofline']['us_per_step'])
print('xyz', d['roofline_knn_xyz'])
print('edgeconv',d['roofline_edgeconv']['frac'], d['roofline_edgeconv']['measured_us'])
for k,v in list(d['kernel_ms_per_step'].items())[:24]: print('  ',k[:100],v)
PY
