#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_lnfold.py 2>&1 | grep -v Warn | tail -14
./tools_gpu_x2.sh
