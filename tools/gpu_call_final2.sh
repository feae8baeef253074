#!/bin/bash
# last check of the committed tree: full -m gpu suite, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/f2_pytest.log; tail -2 gpurun_out/f2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/f2_bench_n1.json 2> gpurun_out/f2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f2_bench_n1.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'knn frac',round(d['roofline']['frac'],3),'edgeconv',round(d['roofline_edgeconv']['frac'],3),'eager ref',round(d['gpu_eager_reference']['value']),'cpu',round(d['cpu_baseline']['value'],1), 'clocks', d['clocks'])
PY
