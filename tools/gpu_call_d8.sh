#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/check_f16.py > gpurun_out/d8_check.log 2>&1; tail -36 gpurun_out/d8_check.log
