"""GEMM microbenchmark: the tcgen05 kernel (several tile widths) vs cuBLAS (torch.matmul) per hot shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
shapes = [(2048, 1280, 1280), (2048, 1280, 5120), (2048, 3840, 1280), (2048, 10240, 1280), (4096, 1280, 1280),
          (4096, 10240, 1280), (4096, 1280, 5120), (8192, 640, 640), (8192, 5120, 640), (8192, 640, 2560),
          (16384, 640, 640), (8192, 8192, 8192)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3  # us (median, cold L2)


for M, N, K in shapes:
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    t_ref = timeit(lambda: torch.matmul(a, w.t(), out=out))
    row = f"M={M:6d} N={N:6d} K={K:5d} cublas {t_ref:7.1f} us {fl / t_ref / 1e6:7.1f} TF/s |"
    best = None
    for bn in (64, 96, 128, 160, 192, 224, 256):
        t = timeit(lambda: ops.gemm(a, w, out, M=M, N=N, K=K, bn=bn))
        row += f" bn{bn}: {t:6.1f}"
        if best is None or t < best[0]:
            best = (t, bn)
    auto = ops.choose_bn(M, N, K)
    row += f" | best bn{best[1]} {fl / best[0] / 1e6:7.1f} TF/s ({best[0] / t_ref:4.2f}x cublas time), model picks bn{auto}"
    print(row, flush=True)
