#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_encoders_gpu.py tests/test_fixtures_gpu.py -q --timeout 900 -rf -x 2>&1 | tee gpurun_out/r02_enc_pytest.log | tail -30
