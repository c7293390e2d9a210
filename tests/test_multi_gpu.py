"""N > 1 on real GPUs (skipped with fewer than 2): CFG-parallel (one all-gather of eps per step over
NCCL) must reproduce the single-GPU CFG run, and data-parallel sharding must reproduce the matching
slice of the batched run.  The CPU (gloo) coverage of the same logic is in test_host_logic.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_cfg_parallel_and_dp_match_single_gpu(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", {"fp32": "29541", "fp16": "29542", "bf16": "29543"}[precision],
                        os.path.join(ROOT, "tools", "cfgp_check.py"), precision],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_CHECK" in r.stdout
