#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[1], {k:d[k] for k in ('value','unit','n_gpus','ms_per_step','gpu_launches','step_tflops','step_frac_of_sustained_peak')}, 'e2e', d['e2e']['value'], d['config']['parallelism'], d['config']['images_per_rank'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 600 python bench.py --workload config3 --no-cpu --no-fp16 --steps 15 > gpurun_out/bench_config3_n1.json 2> gpurun_out/c3.err; echo "c3 rc=$?"; tail -2 gpurun_out/c3.err; show gpurun_out/bench_config3_n1.json
timeout 900 python bench.py --workload config5 --no-cpu --no-fp16 --steps 6 --warmup 3 > gpurun_out/bench_config5_n1.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/c5.err; show gpurun_out/bench_config5_n1.json
timeout 900 python bench.py --workload config4 --batch 4 --no-cpu --no-fp16 --steps 30 --warmup 3 > gpurun_out/bench_config4_b4_n1.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/c4.err; show gpurun_out/bench_config4_b4_n1.json
timeout 900 python bench.py --batch 4 --no-cpu --no-fp16 --steps 15 --warmup 3 > gpurun_out/bench_config2_b4_n1.json 2> gpurun_out/c2b4.err; echo "c2b4 rc=$?"; tail -2 gpurun_out/c2b4.err; show gpurun_out/bench_config2_b4_n1.json
