#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x --timeout 700 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[1], {k:d[k] for k in ('value','unit','n_gpus','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], d['config']['parallelism'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 3 --no-fp16 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/bench_n2.err; show gpurun_out/bench_n2.json
timeout 600 $T --master-port 29514 bench.py --gpus 2 --steps 15 --warmup 3 --workload config3 --cfg-parallel > gpurun_out/bench_config3_n2_cfgp.json 2> gpurun_out/bench_n2c.err; echo "n2 cfgp rc=$?"; tail -2 gpurun_out/bench_n2c.err; show gpurun_out/bench_config3_n2_cfgp.json
timeout 300 $T --master-port 29515 bench.py --gpus 2 --steps 2 --warmup 3 --impl reference | cut -c1-300
