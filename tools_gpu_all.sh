#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -x 2>&1 | tail -15 > gpurun_out/all_tests.log; echo "tests rc=$?"; tail -3 gpurun_out/all_tests.log
for mode in 0 1; do
IIR_NO_PDL=$mode timeout 1200 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_pdl$mode.json 2> gpurun_out/bench.err; echo "bench NO_PDL=$mode rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl$mode.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')})
print('e2e',d['e2e']['value'], d['e2e']['seconds_per_image_batch']); print('roofline',d['roofline']['achieved'], d['roofline']['frac']); print('clocks',d['clocks'])
PY
tail -3 gpurun_out/bench.err
done
cp gpurun_out/bench_pdl0.json gpurun_out/bench.json
timeout 300 python tools/bench_gemm.py 2>&1 | head -8
