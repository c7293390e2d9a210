#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" --timeout 120 2>&1 | tail -3
for pl in 0 8 4 3 2; do echo "IIR_ATTN_POLY=$pl"; IIR_ATTN_POLY=$pl timeout 300 python tools/bench_attn.py 2>&1 | head -4; done
