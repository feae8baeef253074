#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/check_f16.py time > gpurun_out/d3_check.log 2>&1; tail -9 gpurun_out/d3_check.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/d3_pytest.log; tail -3 gpurun_out/d3_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/d3_bench.json 2> gpurun_out/d3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d3_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',d['roofline']['frac'], d['roofline']['frac_of_3xtf32_roofline'], d['roofline']['us_per_step'])
print('edgeconv',d['roofline_edgeconv']['frac'], d['roofline_edgeconv']['measured_us'])
for k,v in list(d['kernel_ms_per_step'].items())[:40]: print('  ',k[:100],v)
PY
timeout 120 python tools/tc_timeline.py model:2 > gpurun_out/d3_tl_model2.txt 2>&1; tail -30 gpurun_out/d3_tl_model2.txt
timeout 120 python tools/tc_timeline.py model:4 > gpurun_out/d3_tl_model4.txt 2>&1
timeout 120 python tools/tc_timeline.py 32,3,1024,20 > gpurun_out/d3_tl_xyz.txt 2>&1
