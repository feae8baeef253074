#!/bin/bash
# final single-GPU evidence: full -m gpu suite, default bench line (with CPU reference + GPU-eager reference legs),
# config-2 line, kNN micro-benchmarks, ncu launch list and full captures of the shipped top kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/f1_pytest.log; tail -3 gpurun_out/f1_pytest.log
timeout 900 python bench.py > gpurun_out/f1_bench_n1.json 2> gpurun_out/f1_bench_n1.err
timeout 600 python bench.py --steps 10 --warmup 3 --points 2048 --k 40 --batch 32 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/f1_bench_cfg2_n1.json 2> gpurun_out/f1_bench_cfg2_n1.err
timeout 300 python tools/check_f16.py > gpurun_out/f1_check_f16.log 2>&1; tail -12 gpurun_out/f1_check_f16.log
timeout 120 python tools/tc_timeline.py model:2 > gpurun_out/f1_tl_model2.txt 2>&1
python tools/profile_step.py > gpurun_out/f1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f1_launches.csv python tools/profile_step.py > gpurun_out/f1_ncu1.log 2>&1
python tools/profile_step.py fwd > gpurun_out/f1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'edge_gather_kernel|knn_tc' -s 24 -c 12 -o gpurun_out/f1_top python tools/profile_step.py fwd > gpurun_out/f1_ncu2.log 2>&1
tail -1 gpurun_out/f1_ncu1.log gpurun_out/f1_ncu2.log
python - <<'PY'
import json
for f in ('gpurun_out/f1_bench_n1.json','gpurun_out/f1_bench_cfg2_n1.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'knn frac',round(d['roofline']['frac'],3),'edgeconv',round(d['roofline_edgeconv']['frac'],3))
PY
