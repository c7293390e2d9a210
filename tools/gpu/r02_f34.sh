#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_infer_cli.py tests/test_model_parity_gpu.py -q --timeout 900 -rf -k "cli or adastep or full_step or config1" 2>&1 | tee gpurun_out/r02_f34_pytest.log | tail -30
