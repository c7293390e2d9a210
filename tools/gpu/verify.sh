#!/bin/bash
# round-end style validation: gpu tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
