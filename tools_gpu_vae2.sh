#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "softmax" --timeout 300 2>&1 | tail -3
timeout 900 python -m pytest tests/test_vae_gpu.py -q --timeout 600 -k full_size 2>&1 | tail -15
timeout 900 python bench.py --no-cpu --no-fp16 --steps 10 > gpurun_out/bench_vae.json 2> gpurun_out/bench_vae.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_vae.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_vae.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['vae_decode'])
PY
