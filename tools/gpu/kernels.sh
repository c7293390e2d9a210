#!/bin/bash
# Bring-up run: each kernel group in its own process under `timeout` so one hang cannot eat the call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 240 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 --timeout-method=thread -k "$1" > gpurun_out/k_$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/k_$name.log; }
run misc "scheduler or linear_small or upsample or concat or layernorm or groupnorm or direct or no_cpu"
run simt "False or simt"
run gemm_tc "test_gemm_linear and True"
run gemm_epi "(silu_rowvec or geglu or sft) and True"
run conv_tc "(test_conv3x3 and True) or im2col"
run attn_tc "attention and True"
cat gpurun_out/summary.txt
