#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" --timeout 120 2>&1 | tail -3
timeout 300 python tools/bench_attn.py > gpurun_out/attn_micro.txt 2>&1; echo "attn micro rc=$?"; cat gpurun_out/attn_micro.txt
