#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-fp16 --no-vae --profiler-range > gpurun_out/b_plain.json 2> gpurun_out/b_plain.err || { tail -3 gpurun_out/b_plain.err; exit 1; }
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-fp16 --no-vae --profiler-range > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"; wc -l gpurun_out/launches_bench_r01c.csv; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
