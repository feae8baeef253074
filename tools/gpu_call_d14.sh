#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/d14_bench_default.json 2> gpurun_out/d14_bench_default.err ) 2> gpurun_out/d14_time.txt
tail -3 gpurun_out/d14_time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d14_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'steps', d['steps'])
print('roofline',{k:d['roofline'][k] for k in ('frac','frac_of_3xtf32_roofline','us_per_step','peak')})
print('gather',{k:d['roofline_gather'][k] for k in ('frac','traffic','lts_bytes_per_step','l2_gbs')})
print('edgeconv',d['roofline_edgeconv']['frac'])
print('cpu',d['cpu_baseline'])
print('eager',d['gpu_eager_reference'])
PY
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/d14_bench_ref.json 2> gpurun_out/d14_bench_ref.err ) 2>> gpurun_out/d14_time.txt
tail -3 gpurun_out/d14_time.txt; cat gpurun_out/d14_bench_ref.json | cut -c1-700
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
