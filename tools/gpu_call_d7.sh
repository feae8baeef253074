#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/d7_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d7_launches.csv python tools/profile_step.py > gpurun_out/d7_ncu1.log 2>&1
python tools/profile_step.py fwd > gpurun_out/d7_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'edge_gather_kernel|knn_tc' -s 22 -c 11 -o gpurun_out/d7_top python tools/profile_step.py fwd > gpurun_out/d7_ncu2.log 2>&1
tail -2 gpurun_out/d7_ncu1.log gpurun_out/d7_ncu2.log; ls -la gpurun_out/d7*
