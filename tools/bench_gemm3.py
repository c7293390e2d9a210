"""K sweep per tile config (in-graph, rotating cold weights): the slope in K is the main-loop rate, the intercept
the fixed cost (launch, prologue, first loads, exposed epilogue)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
from bench_gemm2 import graph_time, R
torch.set_grad_enabled(False)
dev = "cuda"
M, N = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2048x1280").split("x"))
Ks = [640, 1280, 2560, 5120, 10240]
cfgs = [("cublas", None, None), ("1cta bn128", 128, 1), ("1cta bn160", 160, 1), ("1cta bn256", 256, 1), ("pair bn128", 128, 2), ("pair bn160", 160, 2),
        ("pair bn256", 256, 2)]
res = {}
for K in Ks:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(R)]
    ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for name, bn, cl in cfgs:
        if bn is None:
            t = graph_time(lambda: [torch.matmul(a, w.t(), out=ob) for w in ws])
        else:
            t = graph_time(lambda: [ops.gemm(a, w, ob, M=M, N=N, K=K, bn=bn, cluster=cl) for w in ws])
        res[(name, K)] = t
print(f"M={M} N={N}; us per launch at K = {Ks}; slope between the last two K (us per 1000 K) and the MMA-bound slope")
for name, bn, cl in cfgs:
    ts = [res[(name, K)] for K in Ks]
    slope = (ts[-1] - ts[-2]) / (Ks[-1] - Ks[-2]) * 1000
    line = f"{name:12s} " + " ".join(f"{t:7.1f}" for t in ts) + f"   slope {slope:5.2f}"
    if bn:
        tm = (M + 127) // 128
        tn = (N + bn - 1) // bn
        tiles = tm * tn if cl == 1 else ((tm + 1) // 2) * tn
        units = 148 if cl == 1 else 74
        waves = (tiles + units - 1) // units
        ideal = waves * (1000 / 16) * (bn / 2) / 1.9e3
        line += f"  (MMA-bound {ideal:5.2f}; tiles {tiles} on {units} -> {waves} wave(s))"
    print(line, flush=True)
