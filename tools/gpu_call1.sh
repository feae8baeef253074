#!/bin/bash
# round-2 GPU call 1: full GPU suite, 8-rank equivalence diagnostics on one GPU (gloo), bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py 2>&1 | tail -30 > gpurun_out/c1_pytest.log
for W in 2 4 8; do
  DDP_CHECK_BACKEND=gloo DDP_CHECK_ARBITER=1 DDP_CHECK_VERBOSE=1 DDP_CHECK_B=32 OMP_NUM_THREADS=2 timeout 300 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29511 \
    tools/check_ddp_equivalence.py > gpurun_out/c1_ddp_gloo_w${W}_b32.log 2>&1
done
DDP_CHECK_BACKEND=gloo DDP_CHECK_ARBITER=1 DDP_CHECK_VERBOSE=1 DDP_CHECK_B=16 OMP_NUM_THREADS=2 timeout 300 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29512 \
  tools/check_ddp_equivalence.py > gpurun_out/c1_ddp_gloo_w8_b16.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -5 gpurun_out/c1_bench.err
