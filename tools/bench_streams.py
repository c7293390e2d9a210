"""Would two independent kernel chains (aggregator || UNet down path) overlap usefully?  Two chains of the step's
typical launches (out-proj GEMM + LayerNorm + QKV GEMM + attention) captured in one graph on ONE stream vs on TWO
forked streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
M, C, heads, n = 2048, 1280, 20, 1024
L = 10


def make():
    d = {}
    d["a"] = torch.randn(M, C, device=dev, dtype=torch.bfloat16)
    d["h"] = torch.randn(M, C, device=dev)
    d["wo"] = [torch.randn(C, C, device=dev, dtype=torch.bfloat16) * C ** -0.5 for _ in range(L)]
    d["wqkv"] = [torch.randn(3 * C, C, device=dev, dtype=torch.bfloat16) * C ** -0.5 for _ in range(L)]
    d["bo"] = torch.randn(C, device=dev)
    d["g"], d["b"] = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    d["ln"] = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    d["qkv"] = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    d["o"] = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    return d


def chain(d):
    for i in range(L):
        ops.layernorm(d["h"], d["g"], d["b"], d["ln"], rows=M, C=C)
        ops.gemm(d["ln"], d["wqkv"][i], d["qkv"], M=M, N=3 * C, K=C)
        ops.attention(d["qkv"], 0, 3 * C, [d["qkv"]], [C], [3 * C], [d["qkv"]], [2 * C], [3 * C], [n], [1.0], d["o"], 0, C,
                      B=2, heads=heads, n_q=n, softmax_scale=0.125)
        ops.gemm(d["o"], d["wo"][i], d["h"], M=M, N=C, K=C, bias=d["bo"], residual=d["h"])


d1, d2 = make(), make()
chain(d1); chain(d2); torch.cuda.synchronize()
side = torch.cuda.Stream()


def seq():
    chain(d1); chain(d2)


def par():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        chain(d2)
    chain(d1)
    cur.wait_stream(side)


for name, fn in (("one stream", seq), ("two streams", par)):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {best * 1e3:.0f} us for 2 x {L} blocks ({best * 1e3 / (2 * L * 4):.1f} us per launch)")
