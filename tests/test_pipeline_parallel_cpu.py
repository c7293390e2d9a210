"""world_size-2 gloo test (CPU) of the product pipeline's multi-GPU host logic: CFG-parallel and data-parallel runs of
`InstantIRPipeline.__call__`, driven by the oracle modules (tests/_cpu_loop.py), must reproduce the single-process oracle loop —
including a CFG pair whose ranks arrive with different (or no) generator state.  The 2-GPU counterpart (tests/test_multi_gpu.py)
is skipped on a 1-GPU test box; this one always runs."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg_parallel_and_dp_shard_match_the_oracle_loop_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    child = os.path.join(ROOT, "tests", "cfgp_cpu_child.py")
    procs = [subprocess.Popen([sys.executable, child, str(r), str(port)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, cwd=ROOT)
             for r in range(2)]
    outs = [p.communicate(timeout=900)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    assert all("CPU_PARALLEL_CHECK" in o for o in outs)
