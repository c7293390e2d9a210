"""Is FF1 (N = 8C GEGLU) bound by its epilogue?  Same main loop with: GEGLU epilogue / plain bf16 epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instantir_b200 import ops
from bench_gemm2 import graph_time, R
dev = "cuda"
for M in (2048, 4096):
    C = 1280
    a = torch.randn(M, C, device=dev, dtype=torch.bfloat16)
    ws = [torch.randn(8 * C, C, device=dev, dtype=torch.bfloat16) * C ** -0.5 for _ in range(R)]
    b1 = torch.randn(8 * C, device=dev)
    g = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    full = torch.empty(M, 8 * C, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * 8 * C * C
    t1 = graph_time(lambda: [ops.gemm(a, w, g, M=M, N=8 * C, K=C, bias=b1, pair=ops.PAIR_GEGLU, bn=256, cluster=2) for w in ws])
    t2 = graph_time(lambda: [ops.gemm(a, w, full, M=M, N=8 * C, K=C, bias=b1, bn=256, cluster=2) for w in ws])
    t3 = graph_time(lambda: [ops.gemm(a, w, full, M=M, N=8 * C, K=C, bn=256, cluster=2) for w in ws])
    t4 = graph_time(lambda: [torch.matmul(a, w.t(), out=full) for w in ws])
    print(f"M={M}: GEGLU {t1:6.1f} us ({fl / t1 / 1e6:5.0f} TF/s) | plain+bias {t2:6.1f} | plain {t3:6.1f} | cuBLAS {t4:6.1f} ({fl / t4 / 1e6:5.0f} TF/s)")
