#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" --timeout 300 2>&1 | tail -6
timeout 300 python tools/bench_attn.py 2>&1 | grep -v Warn | tail -8
