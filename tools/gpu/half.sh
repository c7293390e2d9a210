#!/bin/bash
mkdir -p gpurun_out
IIR_GEMM_HALF=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv3x3 or folded" --timeout 300 2>&1 | tail -6
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
kb=d.get('kernel_breakdown') or {}
print(sys.argv[1], {k:round(d[k],4) for k in ('value','ms_per_step','step_frac_of_sustained_peak')}, d['clocks']['sm_mhz'], 'single-stream gemm/conv/attn ms:', [round(kb[k]['ms'],2) for k in ('gemm_tc','conv3x3_tc','attn_tc') if k in kb])
PY
}
B="timeout 600 python bench.py --no-cpu --no-fp16 --no-vae --steps 30"
$B > gpurun_out/h_base.json 2> gpurun_out/h.err; show gpurun_out/h_base.json
IIR_GEMM_HALF=1 $B > gpurun_out/h_half.json 2> gpurun_out/h.err; tail -1 gpurun_out/h.err; show gpurun_out/h_half.json
IIR_GEMM_HALF=1 $B --agg-ahead > gpurun_out/h_half_ahead.json 2> gpurun_out/h.err; tail -1 gpurun_out/h.err; show gpurun_out/h_half_ahead.json
