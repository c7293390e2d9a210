#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py config2 > gpurun_out/shapes_config2.txt 2>&1; echo "shapes rc=$?"
head -70 gpurun_out/shapes_config2.txt
# ncu launch list of the bench command (plain run first, same args)
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches_r01.csv
