#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "softmax" --timeout 300 2>&1 | tail -5
timeout 900 python -m pytest tests/test_vae_gpu.py -q --timeout 600 2>&1 | tail -30
