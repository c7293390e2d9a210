#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python tools/autotune.py > gpurun_out/autotune.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/autotune.log
