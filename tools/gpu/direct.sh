#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_vae_gpu.py -q -x --timeout 600 2>&1 | tail -3
timeout 300 python tools/bench_lnfold.py 2>&1 | grep -v Warn | tail -12
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
kb=d['kernel_breakdown']
print(sys.argv[1], {k:round(d[k],4) for k in ('value','ms_per_step','step_frac_of_sustained_peak')}, d['clocks']['sm_mhz'], 'single-stream gemm/conv/attn ms:', [round(kb[k]['ms'],2) for k in ('gemm_tc','conv3x3_tc','attn_tc')], 'roofline', round(d['roofline']['frac'],4))
PY
}
B="timeout 600 python bench.py --no-cpu --no-fp16 --no-vae"
$B > gpurun_out/d_direct.json 2> gpurun_out/d.err; tail -1 gpurun_out/d.err; show gpurun_out/d_direct.json
IIR_GEMM_DIRECT=0 $B > gpurun_out/d_staged.json 2> gpurun_out/d.err; tail -1 gpurun_out/d.err; show gpurun_out/d_staged.json
