#!/bin/bash
timeout 300 python tools/check_f16.py > gpurun_out/d12_check.log 2>&1; tail -14 gpurun_out/d12_check.log
bash tools/gpu_call_d6.sh
