#!/bin/bash
bash tools/gpu_call_d6.sh
bash tools/gpu_call_d15.sh
