#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 300 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_parity_gpu.py tests/test_vae_gpu.py -q -x --timeout 600 2>&1 | tail -3
echo "--- default (16-bit outputs direct, 256-bit stores)"
timeout 300 python tools/bench_lnfold.py 2>&1 | grep -v Warn | tail -12
echo "--- IIR_GEMM_DIRECT=2 (fp32 outputs direct too)"
IIR_GEMM_DIRECT=2 timeout 300 python tools/bench_lnfold.py 2>&1 | grep -v Warn | grep producer
IIR_GEMM_DIRECT=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm or conv3x3 or folded" --timeout 300 2>&1 | tail -2
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
kb=d['kernel_breakdown']
print(sys.argv[1], {k:round(d[k],4) for k in ('value','ms_per_step','step_frac_of_sustained_peak')}, d['clocks']['sm_mhz'], 'single-stream gemm/conv/attn ms:', [round(kb[k]['ms'],2) for k in ('gemm_tc','conv3x3_tc','attn_tc')], 'roofline', round(d['roofline']['frac'],4))
PY
}
B="timeout 600 python bench.py --no-cpu --no-fp16 --no-vae"
$B > gpurun_out/d_direct.json 2> gpurun_out/d.err; tail -1 gpurun_out/d.err; show gpurun_out/d_direct.json
IIR_GEMM_DIRECT=2 $B > gpurun_out/d_direct2.json 2> gpurun_out/d.err; tail -1 gpurun_out/d.err; show gpurun_out/d_direct2.json
