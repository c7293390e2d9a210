"""In-graph GEMM microbenchmark: R launches on rotating weight copies captured in one CUDA graph (no host launch
floor, weights cold like in the real step, activations L2-warm), per hot linear shape of the 1024² step:
plain bf16 output vs the fp32 residual-stream epilogue, several tile configs, and cuBLAS for reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
R = 16
shapes = [(2048, 1280, 1280), (4096, 1280, 1280), (8192, 640, 640), (2048, 1280, 5120), (4096, 1280, 5120),
          (2048, 3840, 1280), (8192, 640, 2560)]
if len(sys.argv) > 1 and __name__ == "__main__":
    shapes = [tuple(int(v) for v in sys.argv[1].split("x"))]


def graph_time(fn, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / R  # us per launch


def main():
    for M, N, K in shapes:
        a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(R)]
        ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        of = torch.randn(M, N, device=dev)
        bias = torch.randn(N, device=dev)
        fl = 2.0 * M * N * K
        t_ref = graph_time(lambda: [torch.matmul(a, w.t(), out=ob) for w in ws])
        print(f"M={M} N={N} K={K}: cuBLAS bf16-out {t_ref:6.1f} us ({fl / t_ref / 1e6:6.0f} TF/s)", flush=True)
        for cl in (1, 2):
            row = f"   cluster={cl}:"
            for bn in (96, 128, 160, 192, 224, 256):
                try:
                    t1 = graph_time(lambda: [ops.gemm(a, w, ob, M=M, N=N, K=K, bias=bias, bn=bn, cluster=cl) for w in ws])
                    t2 = graph_time(lambda: [ops.gemm(a, w, of, M=M, N=N, K=K, bias=bias, residual=of, bn=bn, cluster=cl) for w in ws])
                    row += f"  bn{bn}: {t1:5.1f}/{t2:5.1f}"
                except Exception as e:
                    row += f"  bn{bn}: err"
            print(row + "   (bf16 out / fp32 out + residual, us)", flush=True)



if __name__ == "__main__":
    main()
