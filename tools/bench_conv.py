"""implicit-GEMM conv microbenchmark: 1-CTA vs CTA-pair per hot conv shape (env IIR_GEMM_CLUSTER=1 / 22)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(2, 128, 128, 320, 320), (2, 256, 128, 320, 320), (2, 64, 64, 640, 640), (2, 128, 64, 640, 640), (2, 32, 32, 1280, 1280),
          (2, 64, 32, 1280, 1280), (2, 32, 32, 2560, 1280), (2, 64, 64, 1920, 640), (2, 128, 128, 960, 320), (2, 128, 128, 640, 320),
          (2, 64, 64, 1280, 1280), (2, 64, 64, 128, 1280), (2, 128, 128, 128, 640)]
for n, H, W, Ci, Co in shapes:
    x = torch.randn(n, H, W, Ci, device=dev, dtype=torch.bfloat16)
    w = torch.randn(Co, 9 * Ci, device=dev, dtype=torch.bfloat16)
    M = n * H * W
    out = torch.empty(M, Co, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(x, w, out, M=M, N=Co, K=9 * Ci, conv=dict(n_img=n, H=H, W=W, Cin=Ci))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2] * 1e3
    print(f"conv n={n} {H}x{W} Cin={Ci} Cout={Co}: {t:7.1f} us {2.0 * M * Co * 9 * Ci / t / 1e6:7.1f} TF/s bn={ops.choose_bn(M, Co, 9 * Ci)}", flush=True)
