#!/bin/bash
mkdir -p gpurun_out
python tools/profile_knn.py 32,64,1024,20 > gpurun_out/d4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 4 -c 1 -o gpurun_out/d4_knn16_c64 python tools/profile_knn.py 32,64,1024,20 > gpurun_out/d4_ncu.log 2>&1
tail -3 gpurun_out/d4_ncu.log
ls -la gpurun_out/*.ncu-rep
