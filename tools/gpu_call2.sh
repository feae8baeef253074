#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/c2_pytest.log
DDP_CHECK_BACKEND=gloo DDP_CHECK_ARBITER=1 DDP_CHECK_SEED=5 DDP_CHECK_B=32 OMP_NUM_THREADS=2 timeout 300 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29511 \
  tools/check_ddp_equivalence.py > gpurun_out/c2_ddp_gloo_w8_b32_seed5.log 2>&1
python tools/tc_timeline.py > gpurun_out/c2_tc_timeline.txt 2>&1
tail -3 gpurun_out/c2_pytest.log
