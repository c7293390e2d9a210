"""CPU test of bench.py's control flow (the part no GPU-less box can otherwise run): `run_ours` with the device layer stubbed —
CUDA queries, events, models and the pipeline are replaced by stand-ins that keep tensors on the CPU.  Checks that the single
JSON line is assembled (base_line + roofline leg), that the opt-in experimental leg is a contained child-process call, and that
the partition-phase deadline prints the line collected so far and leaves (bench.py's safety net for a one-sided failure inside a
CFG pair)."""
import json
import os
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 12.5


class _Loop:
    def __init__(self):
        self.latents = torch.arange(4 * 16 * 16, dtype=torch.float32).view(1, 4, 16, 16) / 100.0
        self.n_steps = 30
        self.steps = 0

    def step(self, i):
        self.steps += 1
        self.latents = self.latents * 0.99 + 0.01
        return self.latents


class _Pipe:
    def __init__(self, *a, **k):
        self._graphs = {}

    def __call__(self, **kw):
        if kw.get("prepare_only"):
            return _Loop()
        return types.SimpleNamespace(images=torch.zeros(1, 4, 16, 16))


@pytest.fixture
def stubbed_bench(monkeypatch):
    import bench
    from instantir_b200 import pipeline

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch, "Generator", lambda device=None: types.SimpleNamespace(manual_seed=lambda s: None))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a)
        k.pop("non_blocking", None)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", to_cpu)
    monkeypatch.setattr(bench, "build_models", lambda *a, **k: (object(), object()))
    monkeypatch.setattr(bench, "host_inputs", lambda cfg, B, latent, seed=0: {"image": torch.zeros(B, 4, 16, 16)})
    monkeypatch.setattr(pipeline, "InstantIRPipeline", _Pipe)
    monkeypatch.setattr(pipeline.LaunchCounter, "total", classmethod(lambda cls: 0))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1800.0, "sm_max_mhz": 1965.0, "reasons": []})
    lines = []
    monkeypatch.setattr(bench, "emit", lambda line: lines.append(json.loads(json.dumps(line))))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return bench, lines


def _args(bench, *extra):
    argv = ["bench.py", "--steps", "4", "--warmup", "3", "--no-cpu", "--no-alt", "--no-vae", *extra]
    old = sys.argv
    sys.argv = argv
    try:
        # the parser main() builds: reuse it by calling main() with run_ours captured
        captured = {}
        real = bench.run_ours
        bench.run_ours = lambda a: captured.setdefault("args", a)
        guard = bench._guard_stdout
        bench._guard_stdout = lambda: None
        try:
            bench.main()
        finally:
            bench.run_ours, bench._guard_stdout = real, guard
        return captured["args"]
    finally:
        sys.argv = old


def test_single_gpu_line_is_assembled_with_contained_experimental_leg(stubbed_bench, monkeypatch):
    bench, lines = stubbed_bench
    child_line = {"ms_per_step": 3.0, "value": 11.0, "unit": "img/s", "gpu_launches": 7,
                  "kernel_breakdown": {"groupnorm": {"ms": 0.2}, "groupnorm_apply": {"ms": 0.5}}, "latents_probe_after_step0": None}
    calls = []

    def fake_run(cmd, **kw):
        calls.append((cmd, kw))
        return types.SimpleNamespace(returncode=0, stdout="noise\n" + json.dumps(child_line) + "\n", stderr="")

    monkeypatch.setattr(subprocess, "run", fake_run)
    bench.run_ours(_args(bench))
    assert len(lines) == 1
    line = lines[0]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "partitions", "kernel_breakdown"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["dtype"] == "fp16" and line["partitions"] is None
    assert line["ms_per_step"] == pytest.approx(12.5 / 4) and line["value"] == pytest.approx(1 / (30 * 12.5 / 4 * 1e-3))
    probe = line["latents_probe_after_step0"]
    assert probe["norm"] > 0 and len(probe["sample"]) == 64
    exp = line["experimental_gn_fuse"]
    assert exp["status"] == "ran" and exp["ms_per_step"] == 3.0 and exp["norm_kernels_single_stream_ms"]["fused"]["groupnorm_apply"] == 0.5
    (cmd, kw), = calls
    assert "--no-experimental" in cmd and kw["env"]["IIR_GN_FUSE"] == "1" and kw["timeout"] <= 300
    # a crashing child is a recorded failure, not an exception
    monkeypatch.setattr(subprocess, "run", lambda cmd, **kw: types.SimpleNamespace(returncode=-11, stdout="", stderr="segfault"))
    lines.clear()
    bench.run_ours(_args(bench))
    assert lines[0]["experimental_gn_fuse"]["status"] == "failed" and lines[0]["value"] > 0

    def boom(cmd, **kw):
        raise subprocess.TimeoutExpired(cmd, kw["timeout"])

    monkeypatch.setattr(subprocess, "run", boom)
    lines.clear()
    bench.run_ours(_args(bench))
    assert lines[0]["experimental_gn_fuse"]["status"] == "failed" and "TimeoutExpired" in lines[0]["experimental_gn_fuse"]["error"]
    lines.clear()
    bench.run_ours(_args(bench, "--no-experimental"))
    assert "experimental_gn_fuse" not in lines[0]


def test_partition_deadline_prints_the_line_and_leaves(stubbed_bench, monkeypatch):
    """2 ranks (stubbed collectives), the first partition sub-run never returns: the deadline thread must emit the main line with
    the partitions collected so far and call os._exit(0)"""
    import threading
    import time

    import torch.distributed as dist

    bench, lines = stubbed_bench
    from instantir_b200 import parallel

    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda: None)
    monkeypatch.setattr(parallel, "CFGParallel", lambda: types.SimpleNamespace(branch=0))
    monkeypatch.setattr(torch, "tensor", lambda data, device=None, dtype=None: torch.as_tensor(data, dtype=dtype))
    exited = threading.Event()
    monkeypatch.setattr(os, "_exit", lambda code: exited.set())
    real_timer = threading.Timer
    monkeypatch.setattr(threading, "Timer", lambda secs, fn: real_timer(0.3, fn))  # the deadline, shortened
    hang = threading.Event()
    n_prepared = []

    class _HangingPipe(_Pipe):
        def __call__(self, **kw):
            if kw.get("prepare_only"):
                n_prepared.append(1)
                if len(n_prepared) == 2:  # the first partition sub-run: wait "in a collective" until the test lets go
                    hang.wait(5.0)
                    raise RuntimeError("released by the test")
            return super().__call__(**kw)

    from instantir_b200 import pipeline

    monkeypatch.setattr(pipeline, "InstantIRPipeline", _HangingPipe)
    t = threading.Thread(target=lambda: bench.run_ours(_args(bench)), daemon=True)
    t.start()
    assert exited.wait(4.0), "the partition deadline did not fire"
    hang.set()
    t.join(5.0)
    assert lines, "no line was emitted at the deadline"
    first = lines[0]
    assert "partitions_aborted" in first and first["partitions"] == {} and first["value"] > 0 and first["n_gpus"] == 2
    time.sleep(0.05)


def test_two_rank_run_records_the_communicating_partitions(stubbed_bench, monkeypatch):
    """WORLD_SIZE=2 with stubbed collectives: the main line carries the config-3 CFG-parallel and data-parallel sub-records and
    the deadline timer is cancelled"""
    import threading

    import torch.distributed as dist

    bench, lines = stubbed_bench
    from instantir_b200 import parallel

    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("LOCAL_RANK", "0")
    for name in ("init_process_group", "barrier", "destroy_process_group"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)
    monkeypatch.setattr(dist, "all_reduce", lambda t, op=None: None)
    monkeypatch.setattr(parallel, "CFGParallel", lambda: types.SimpleNamespace(branch=0))
    monkeypatch.setattr(torch, "tensor", lambda data, device=None, dtype=None: torch.as_tensor(data, dtype=dtype))
    timers = []
    real_timer = threading.Timer

    def make_timer(secs, fn):
        t = real_timer(secs, fn)
        timers.append((secs, t))
        return t

    monkeypatch.setattr(threading, "Timer", make_timer)
    monkeypatch.setattr(os, "_exit", lambda code: (_ for _ in ()).throw(AssertionError("the deadline fired in a healthy run")))
    bench.run_ours(_args(bench))
    (secs, t), = timers
    assert secs >= 90.0 and t.finished.is_set() or not t.is_alive()  # cancelled
    line, = lines
    assert line["n_gpus"] == 2 and "partitions_aborted" not in line
    parts = line["partitions"]
    assert set(parts) == {"config3_cfg_parallel", "config3_data_parallel"}
    for rec in parts.values():
        assert rec["value"] > 0 and rec["ms_per_step"] > 0 and "error" not in rec
    assert parts["config3_cfg_parallel"]["images"] == 1 and parts["config3_data_parallel"]["images"] == 2
    assert "experimental_gn_fuse" not in line  # N = 1 only
