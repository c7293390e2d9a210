#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
cd tools
timeout 600 python bench_gemm3.py 2048x1280 > ../gpurun_out/gemm_ksweep.txt 2>&1; echo "rc=$?"
timeout 600 python bench_gemm3.py 4096x1280 >> ../gpurun_out/gemm_ksweep.txt 2>&1; echo "rc=$?"
cat ../gpurun_out/gemm_ksweep.txt
