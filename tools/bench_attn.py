"""Attention microbenchmark: attn_tc_kernel per hot shape of the 1024² step vs torch SDPA (library reference)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from instantir_b200 import ops
torch.set_grad_enabled(False)
dev = "cuda"
H16 = H16 if os.environ.get("IIR_TEST_H16", "fp16") == "bf16" else torch.float16
only = sys.argv[1] if len(sys.argv) > 1 else None
# (B, heads, n_q, kv_lens)
shapes = [(2, 20, 1024, [1024]), (2, 20, 2048, [2048]), (2, 10, 4096, [4096]), (2, 10, 8192, [8192]),
          (1, 20, 1024, [1024]), (1, 10, 4096, [4096]), (2, 10, 16384, [16384]),
          (2, 20, 1024, [77, 64]), (2, 10, 4096, [77, 64])]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
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
    return ts[len(ts) // 2] * 1e3


for B, heads, n, kvs in shapes:
    if only and str(n) != only:
        continue
    C = heads * 64
    q = torch.randn(B, n, C, device=dev, dtype=H16)
    ks = [torch.randn(B, m, C, device=dev, dtype=H16) for m in kvs]
    vs = [torch.randn(B, m, C, device=dev, dtype=H16) for m in kvs]
    out = torch.empty(B, n, C, device=dev, dtype=H16)
    fl = 4.0 * B * heads * n * sum(kvs) * 64
    t = timeit(lambda: ops.attention(q, 0, C, ks, [0] * len(kvs), [C] * len(kvs), vs, [0] * len(kvs), [C] * len(kvs), kvs,
                                     [1.0] * len(kvs), out, 0, C, B=B, heads=heads, n_q=n, softmax_scale=0.125))
    qh = q.view(B, n, heads, 64).transpose(1, 2)
    def ref():
        o = None
        for k, v in zip(ks, vs):
            r = F.scaled_dot_product_attention(qh, k.view(B, -1, heads, 64).transpose(1, 2), v.view(B, -1, heads, 64).transpose(1, 2))
            o = r if o is None else o + r
        return o
    t_ref = timeit(ref)
    r = ref().transpose(1, 2).reshape(B, n, C).float()
    err = ((out.float() - r).norm() / r.norm()).item()
    print(f"B={B} heads={heads} n_q={n} kv={kvs}: ours {t:7.1f} us {fl / t / 1e6:7.1f} TF/s | torch SDPA {t_ref:7.1f} us {fl / t_ref / 1e6:7.1f} TF/s | rel err vs SDPA {err:.2e}", flush=True)
