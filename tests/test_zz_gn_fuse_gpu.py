"""OPT-IN fused GroupNorm path (IIR_GN_FUSE=1): GroupNorm statistics accumulated by the epilogue of the GEMM /
implicit-GEMM conv that produces the tensor (`iir_gemm_args.gn_sums`, int64 fixed point) + the one-pass
`iir_groupnorm_apply_sums` — the north star's "GroupNorm fused into the conv epilogue" (SURVEY §2.2 K2/K10,
module/min_sdxl.py:245,250,568,838).

This path was written after the round's GPU budget was spent and has NEVER run on a GPU: it is off by default, the default
kernels' SASS is unchanged (checked with cuobjdump), and each test below runs in a CHILD process (tests/gn_fuse_child.py) with a
timeout, so that neither a fault nor a hang in it can touch the rest of the suite.  The tests are non-strict xfails: an XPASS
in the driver's log is the verification, an XFAIL says the path is still broken (the child's error is printed).  The file name
sorts last on purpose: the verified suite has finished before the first unverified kernel is launched."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNVERIFIED = ("opt-in path written without GPU access (round 2, GPU budget spent): never run before the driver's own GPU test run; "
              "isolated in a child process; XPASS = verified")


def _child(what, timeout):
    env = dict(os.environ)
    env.pop("IIR_GN_FUSE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gn_fuse_child.py"), what], capture_output=True, text=True,
                       cwd=ROOT, timeout=timeout, env=env)
    lines = [ln for ln in r.stdout.strip().split("\n") if ln.startswith("{")]
    assert lines, f"no result line (rc={r.returncode}): {r.stdout[-800:]} {r.stderr[-1500:]}"
    res = json.loads(lines[-1])
    for k, v in res["checks"].items():
        print(f"  {k}: {v}")
    assert res["ok"], res["error"]
    return res["checks"]


@pytest.mark.xfail(reason=UNVERIFIED, strict=False)
def test_gn_sums_and_apply_kernels_in_child_process():
    """gemm(gn=...) on linear and conv launches (several tilings, CTA pairs, ragged N tiles, 16-bit and fp32 + residual
    outputs): fixed-point sums vs fp64 torch sums, run-to-run bit identity, rejection of an ineligible launch;
    groupnorm_apply_sums vs torch GroupNorm and vs the two-kernel iir_groupnorm."""
    _child("kernels", 300)


@pytest.mark.xfail(reason=UNVERIFIED, strict=False)
def test_config1_full_step_with_fused_groupnorm_in_child_process():
    """BASELINE config 1, 2 steps, CFG 7, previewer on, fp16, eager and CUDA-graph: per-step latents of the fused path vs
    the CPU oracle <= 1e-2 (the same bar as the default path), and fewer launches than the default path."""
    _child("model", 420)


@pytest.mark.xfail(reason=UNVERIFIED, strict=False)
def test_sdxl_width_step_fused_vs_default_in_child_process():
    """one UNet + Aggregator step at full SDXL widths (latent 32²): fused vs default path on identical weights < 2e-3"""
    _child("sdxl", 480)
