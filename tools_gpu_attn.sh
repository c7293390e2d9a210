#!/bin/bash
mkdir -p gpurun_out
python tools/bench_attn.py > gpurun_out/attn_micro.txt 2>&1; echo "attn micro rc=$?"; cat gpurun_out/attn_micro.txt
python tools/bench_attn.py 4096 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 4 -c 1 -o gpurun_out/prof_attn python tools/bench_attn.py 4096 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
