#!/bin/bash
# quick step check: attention tests + model parity + bench with per-shape breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout 600 -rf 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_parity_gpu.py -q --timeout 600 -rf -k "not sdxl_width and not full_size" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-alt --no-vae > gpurun_out/r02_bench_step.json 2> gpurun_out/r02_bench_step.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_step.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_step.json').read().strip().split('\n')[-1])
kb=d.pop('kernel_breakdown')
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_frac_of_sustained_peak','dtype')}, d['clocks'])
for k,v in sorted(kb.items(), key=lambda kv:-kv[1]['ms'])[:8]: print(f"{k:18s} n={v['launches']:4d} ms={v['ms']:.3f} tflops={v['tflops']} gbs={v['gbs']}")
for k,v in d['top_shapes'].items(): print(f"{k:50s} n={v['launches']:4d} ms={v['ms']:.3f} us={v['us_per_launch']:.1f} tflops={v['tflops']:.0f}")
PY
