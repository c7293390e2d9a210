#!/bin/bash
# attention split-range scheduling: kernel tests (default + forced split), microbench per mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_kernels_bf16_gpu.py -q -k "attention" --timeout 600 -rf 2>&1 | tee gpurun_out/r02_attn_pytest.log | tail -8
IIR_ATTN_SPLIT=2 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout 600 -rf 2>&1 | tee gpurun_out/r02_attn_pytest_split2.log | tail -8
for m in 0 1 2; do echo "== IIR_ATTN_SPLIT=$m"; IIR_ATTN_SPLIT=$m timeout 300 python tools/bench_attn.py 2>&1 | tee gpurun_out/r02_attn_split$m.txt; done
