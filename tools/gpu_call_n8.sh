#!/bin/bash
# 8-GPU box: multi-rank parity (NCCL + peer exchange at 2/4/8 ranks), weak + strong scaling, config 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "nccl" 2>&1 | tail -4 > gpurun_out/n8_pytest.log; tail -2 gpurun_out/n8_pytest.log
run() { # name N extra...
  name=$1; N=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${name}.json').read().strip().splitlines()[-1])
    print('${name}: value %.0f clouds/s  %.3f ms/step  e2e %.0f  scaling %s  B/GPU %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['scaling'], d['config']['global_batch']//d['n_gpus']))
except Exception as e:
    print('${name}: FAILED', e)
PY
}
run n8_weak 8
run n2_weak 2
run n4_weak 4
run n8_strong 8 --scaling strong
run n8_cfg2_strong 8 --scaling strong --points 2048 --k 40
# single-GPU line of the same build on the same box, for the ratio
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-reference > gpurun_out/n1_same_box.json 2> gpurun_out/n1_same_box.err
python -c "
import json
d=json.loads(open('gpurun_out/n1_same_box.json').read().strip().splitlines()[-1]); print('n1_same_box: value %.0f  %.3f ms/step' % (d['value'], d['ms_per_step']))"
