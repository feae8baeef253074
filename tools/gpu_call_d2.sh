#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_f16.py > gpurun_out/d2_check.log 2>&1; tail -12 gpurun_out/d2_check.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/d2_pytest.log; tail -5 gpurun_out/d2_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/d2_bench.json 2> gpurun_out/d2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d2_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',d['roofline'])
print('edgeconv',d['roofline_edgeconv'])
for k,v in list(d['kernel_ms_per_step'].items())[:30]: print('  ',k[:100],v)
PY
