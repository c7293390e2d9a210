"""Per-kernel-class breakdown of one SDXL VAE decode / encode at 1024² (eager launches, CUDA events around every op)."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instantir_b200 import ops, weights
from instantir_b200.vae import AutoencoderKL, VaeConfig, vae_param_shapes
torch.set_grad_enabled(False)
dev = "cuda"
cfg = VaeConfig()
vae = AutoencoderKL(cfg, weights.RandomSource(vae_param_shapes(cfg), dev, seed=2), dev, "bf16")
z = torch.randn(1, 4, 128, 128, device=dev)
x = torch.rand(1, 3, 1024, 1024, device=dev) * 2 - 1
for name, fn in (("decode", lambda: vae.decode(z).sample), ("encode", lambda: vae.encode(x).latent_dist.mode())):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ops.PROFILE = []
    fn()
    torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for n, w, a, b in ops.PROFILE:
        agg[n][0] += 1; agg[n][1] += a.elapsed_time(b); agg[n][2] += w.get("flops", 0.0)
    ops.PROFILE = None
    tot = sum(v[1] for v in agg.values())
    print(f"{name}: {e0.elapsed_time(e1):.2f} ms wall (eager), sum of op times {tot:.2f} ms, {sum(v[0] for v in agg.values())} launches")
    for k, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:16s} n={n:4d} {ms:7.3f} ms" + (f"  {fl / ms / 1e9:7.0f} TFLOP/s" if fl else ""))
