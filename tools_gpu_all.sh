#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -6 gpurun_out/all_tests.log
timeout 600 python tools/diag_bf16.py 2>&1 | tail -12
