#!/bin/bash
# 2-GPU box: multi-rank parity (NCCL + peer exchange, 2 ranks), weak-scaling line, single-GPU line of the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "nccl" 2>&1 | tail -4 > gpurun_out/n2_pytest.log; tail -2 gpurun_out/n2_pytest.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "shared_gpu" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/n2_weak.json 2> gpurun_out/n2_weak.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/n1_same_box2.json 2> gpurun_out/n1_same_box2.err
python - <<'PY'
import json
for f in ('n2_weak','n1_same_box2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print('%s: value %.0f  %.3f ms/step' % (f, d['value'], d['ms_per_step']))
PY
