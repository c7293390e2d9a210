#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
timeout 600 python tools/bench_gemm2.py > gpurun_out/gemm_micro2.txt 2>&1; echo "rc=$?"; cat gpurun_out/gemm_micro2.txt
