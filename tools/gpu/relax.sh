#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 2>&1 | tail -3
timeout 300 python tools/bench_lnfold.py 2>&1 | grep -v Warn | tail -12
timeout 300 python tools/probe_epilogue.py 2>&1 | grep -v Warn | tail -15
