#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/cfgp_check.py fp32 2>&1 | grep -v "^W\|^\[W" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/cfgp_check.py bf16 2>&1 | grep -v "^W\|^\[W" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['config']['parallelism'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --cfg-parallel > gpurun_out/bench_n2_cfgp.json 2> gpurun_out/bench_n2.err; echo "bench n2 cfgp rc=$?"; tail -3 gpurun_out/bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_cfgp.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], d['config']['parallelism'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 2 --warmup 3 --impl reference | cut -c1-400
