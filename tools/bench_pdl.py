"""Does Programmatic Dependent Launch survive CUDA-graph capture here?  A chain of tiny library kernels (silu on a
few KB) is captured with and without PDL (IIR_NO_PDL=1 in a second process) and replayed."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) == 1:
    for env in ({}, {"IIR_NO_PDL": "1"}):
        e = dict(os.environ); e.update(env)
        print(env or "PDL on", subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True).stdout.strip())
    sys.exit(0)
import torch
from instantir_b200 import ops
dev = "cuda"
N = 400
for numel, rows in ((4096, 0), (1 << 20, 0), (0, 2048)):
    if rows:
        x = torch.randn(rows, 1280, device=dev); g = torch.ones(1280, device=dev); b = torch.zeros(1280, device=dev)
        y = torch.empty(rows, 1280, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.layernorm(x, g, b, y, rows=rows, C=1280)
        tag = f"layernorm {rows}x1280"
    else:
        x = torch.randn(numel, device=dev); y = torch.empty_like(x)
        fn = lambda: ops.silu(x, y)
        tag = f"silu {numel}"
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(N):
            fn()
    g_.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) * 1e3 / N:.2f} us/kernel;", end=" ")
print()
