"""In-graph cost of the folded-LayerNorm epilogues (producer: + 16-bit copy + row partial sums; consumer: + mean /
rstd correction) against the plain GEMMs and the LayerNorm launch they replace, on the step's hot shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
from tools.bench_gemm2 import graph_time, R
torch.set_grad_enabled(False)
dev = "cuda"


class St:
    def __init__(self, M, C):
        self.h16 = torch.randn(M, C, device=dev).to(torch.bfloat16)
        self.acc = torch.zeros(2, M, 2, device=dev, dtype=torch.int64)
        self.cur = 0


for M, C in ((2048, 1280), (4096, 1280), (8192, 640)):
    a = torch.randn(M, C, device=dev, dtype=torch.bfloat16)
    ws = [torch.randn(C, C, device=dev, dtype=torch.bfloat16) * C ** -0.5 for _ in range(R)]
    h = torch.randn(M, C, device=dev)
    bias = torch.randn(C, device=dev)
    st = St(M, C)
    t0 = graph_time(lambda: [ops.gemm(a, w, h, M=M, N=C, K=C, bias=bias, residual=h) for w in ws])
    t1 = graph_time(lambda: [ops.gemm(a, w, h, M=M, N=C, K=C, bias=bias, residual=h, ln_out=st) for w in ws])
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    o16 = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    tl = graph_time(lambda: [ops.layernorm(h, g, b, o16, rows=M, C=C, eps=1e-5) for _ in ws])
    print(f"M={M} C={C}: producer (out-proj + residual) plain {t0:5.1f} us, + ln_out {t1:5.1f} us; layernorm launch {tl:5.1f} us", flush=True)
    for N2, pair in ((3 * C, 0), (C, 0), (8 * C, 1)):
        w2 = [torch.randn(N2, C, device=dev, dtype=torch.bfloat16) * C ** -0.5 for _ in range(R)]
        cs = torch.randn(N2, device=dev)
        b2 = torch.randn(N2, device=dev)
        o = torch.empty(M, N2 // 2 if pair else N2, device=dev, dtype=torch.bfloat16)
        kw = dict(pair=ops.PAIR_GEGLU, bn=256) if pair else {}
        c0 = graph_time(lambda: [ops.gemm(a, w, o, M=M, N=N2, K=C, bias=b2, **kw) for w in w2])
        c1 = graph_time(lambda: [ops.gemm(st.h16, w, o, M=M, N=N2, K=C, bias=b2, ln_in=(st, cs, 1e-5), **kw) for w in w2])
        print(f"      consumer N={N2} pair={pair}: plain {c0:5.1f} us, + ln_in {c1:5.1f} us", flush=True)
