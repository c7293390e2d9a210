#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/trace_step.py config2 1 > gpurun_out/trace_config2_now.txt 2>&1; echo "trace rc=$?"
timeout 300 python tools/diag_bf16.py > gpurun_out/diag_plain.txt 2>&1; echo "diag rc=$?"
IIR_DIAG_ROUND_W=1 timeout 300 python tools/diag_bf16.py > gpurun_out/diag_roundw.txt 2>&1; echo "diag2 rc=$?"
grep -v Warn gpurun_out/diag_plain.txt | tail -12; grep -v Warn gpurun_out/diag_roundw.txt | tail -12
