#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_parity_gpu.py -q --timeout 300 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/models.log
echo "models rc=$?"; tail -8 gpurun_out/models.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_first.json; tail -20 gpurun_out/bench_first.err
