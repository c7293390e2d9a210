"""Where does a one-wave GEMM launch spend its time?  Builds a measurement variant of the library
(-DIIR_GEMM_PROBE) whose GEMM epilogue can be switched off (mode 1: arrive right after the accumulator is ready;
mode 2: everything but the global stores) and times the step's small shapes in-graph.  Not a bench number."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instantir_b200 import build as _b
probe = os.path.join(_b.HERE, "libinstantir_b200_probe.so")
_b._build_variant(probe, "probe", ["-DIIR_GEMM_PROBE=1"], False, False, False)
os.environ["IIR_LIB_OVERRIDE"] = probe
import torch
from instantir_b200 import ops
from tools.bench_gemm2 import graph_time, R
torch.set_grad_enabled(False)
dev = "cuda"


class St:
    def __init__(self, M, C):
        self.h16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
        self.acc = torch.zeros(2, M, 2, device=dev, dtype=torch.int64)
        self.cur = 0


for M, N, K in ((2048, 1280, 1280), (4096, 1280, 1280), (2048, 3840, 1280), (4096, 3840, 1280), (8192, 640, 640), (8192, 1920, 640)):
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(R)]
    h = torch.randn(M, N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev)
    st = St(M, N)
    row = f"M={M} N={N} K={K}:"
    for name, fn in (("fp32+res+ln_out", lambda w: ops.gemm(a, w, h, M=M, N=N, K=K, bias=bias, residual=h, ln_out=st)),
                     ("bf16 out", lambda w: ops.gemm(a, w, o16, M=M, N=N, K=K, bias=bias))):
        ts = []
        for mode in (0, 2, 3, 4, 1):
            os.environ["IIR_GEMM_PROBE_MODE"] = str(mode)
            ts.append(graph_time(lambda: [fn(w) for w in ws]))
        row += (f"\n    {name:16s}: full {ts[0]:5.1f} | no global stores {ts[1]:5.1f} | tmem ld + math only {ts[2]:5.1f} | tmem ld only {ts[3]:5.1f} "
                f"| no epilogue {ts[4]:5.1f} us")
    print(row, flush=True)
os.environ["IIR_GEMM_PROBE_MODE"] = "0"
