#!/bin/bash
mkdir -p gpurun_out
IIR_NO_PDL=1 timeout 600 python tools/trace_step.py config2 1 > gpurun_out/trace_config2_nopdl.txt 2>&1; echo "trace rc=$?"; grep -v Warn gpurun_out/trace_config2_nopdl.txt | tail -62
