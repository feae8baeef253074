#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference --profile-ranks > gpurun_out/n2p_weak.json 2> gpurun_out/n2p_weak.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2p_weak.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'sum',d['kernel_ms_per_step_sum'])
for k,v in list(d['kernel_ms_per_step'].items())[:60]: print('  ',k[:120],v)
PY
