#!/bin/bash
# quick regression + bench after a kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 2>&1 | tail -2
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_vae_gpu.py -q -x --timeout 600 2>&1 | tail -2
timeout 600 python bench.py --no-cpu --no-fp16 --no-vae > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -1 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
kb=d['kernel_breakdown']
print({k:round(d[k],4) for k in ('value','ms_per_step','step_frac_of_sustained_peak')}, d['clocks']['sm_mhz'], 'single-stream gemm/conv/attn ms:', [round(kb[k]['ms'],2) for k in ('gemm_tc','conv3x3_tc','attn_tc')], 'roofline', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],4))
PY
