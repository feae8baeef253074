#!/bin/bash
mkdir -p gpurun_out
for CL in 2 1 4; do
  ECB200_KNN_CLUSTER=$CL timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "knn or block_tensor or dgcnn" 2>&1 | tail -15 > gpurun_out/c4_pytest_cl$CL.log
  echo "cluster $CL: $(tail -1 gpurun_out/c4_pytest_cl$CL.log)"
  ECB200_KNN_CLUSTER=$CL timeout 120 python tools/tc_timeline.py 32,64,1024,20 > gpurun_out/c4_timeline_c64_cl$CL.txt 2>&1
  ECB200_KNN_CLUSTER=$CL timeout 120 python tools/tc_timeline.py 32,128,1024,20 > gpurun_out/c4_timeline_c128_cl$CL.txt 2>&1
  tail -3 gpurun_out/c4_timeline_c64_cl$CL.txt; tail -3 gpurun_out/c4_timeline_c128_cl$CL.txt
done
