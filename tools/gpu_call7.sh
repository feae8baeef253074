#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c7_pytest.log; tail -3 gpurun_out/c7_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',d['roofline'])
print('edgeconv',d['roofline_edgeconv'])
for k,v in list(d['kernel_ms_per_step'].items())[:24]: print('  ',k,v)
PY
