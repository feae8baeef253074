#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "nccl" 2>&1 | tail -5
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
print('N=${N} value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['config']['bn_stats_exchange'], 'eager', d['config']['eager_ms_per_step'])
PY
tail -3 gpurun_out/n${N}_bench.err
