#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_parity_gpu.py -q --timeout 300 --timeout-method=thread -x 2>&1 | tail -60 > gpurun_out/models.log
echo "models rc=$?" ; tail -60 gpurun_out/models.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
