"""GroupNorm / LayerNorm microbenchmark, in-graph over rotating buffers: us per op and GB/s of algorithmic traffic."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
R = 8


def graph_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / R


for n, H, W, C in [(2, 128, 128, 320), (2, 256, 128, 320), (2, 64, 64, 640), (2, 128, 64, 640), (2, 32, 32, 1280), (2, 64, 32, 1280),
                   (2, 32, 32, 2560), (2, 64, 64, 1920), (2, 128, 128, 960)]:
    HW = H * W
    g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for dt in (torch.float32, torch.bfloat16):
        xs = [torch.randn(n * HW, C, device=dev).to(dt) for _ in range(R)]
        out = torch.empty(n * HW, C, device=dev, dtype=torch.bfloat16)
        t = graph_time(lambda: [ops.groupnorm(x, g_, b_, out, n_img=n, HW=HW, C=C, silu=True) for x in xs])
        by = n * HW * C * (2 * xs[0].element_size() + 2)
        print(f"groupnorm n={n} {H}x{W} C={C} {str(dt)[6:]:8s}: {t:6.1f} us  {by / t / 1e3:6.0f} GB/s", flush=True)
for rows, C in [(2048, 1280), (4096, 1280), (8192, 640), (16384, 640)]:
    g_, b_ = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    xs = [torch.randn(rows, C, device=dev) for _ in range(R)]
    out = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
    t = graph_time(lambda: [ops.layernorm(x, g_, b_, out, rows=rows, C=C) for x in xs])
    by = rows * C * 6
    print(f"layernorm rows={rows} C={C}: {t:6.1f} us  {by / t / 1e3:6.0f} GB/s", flush=True)
