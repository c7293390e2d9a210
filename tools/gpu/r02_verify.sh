#!/bin/bash
# round 2: gpu tests (all, no -x), smoke, both bench arms; logs under gpurun_out/
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread -rxXf 2>&1 | tee gpurun_out/r02_pytest.log | tail -40
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().split('\n')[-1])
kb=d.pop('kernel_breakdown')
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','step_frac_of_sustained_peak','dtype')}, d['clocks'])
print('e2e',d['e2e']['value'],'roofline',d['roofline']['frac'], 'alt', d.get('bf16_out_of_spec'), 'vae', d.get('vae_decode'), 'cpu', d.get('cpu_baseline'))
for k,v in sorted(kb.items(), key=lambda kv:-kv[1]['ms']): print(f"{k:18s} n={v['launches']:4d} ms={v['ms']:.3f} tflops={v['tflops']} gbs={v['gbs']}")
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_ref.json; tail -2 gpurun_out/r02_bench_ref.err
