"""Per-shape breakdown of one eager config-2 step (CUDA events around every launch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from instantir_b200 import config as pcfg, ops
from instantir_b200.pipeline import InstantIRPipeline
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
torch.set_grad_enabled(False)
dev = "cuda:0"
wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
cfg = pcfg.sdxl()
unet, agg = bench.build_models(cfg, dev, "bf16", with_lora=(wl == "config3"))
pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
host = bench.host_inputs(cfg, 1, 128)
devin = {k: v.to(dev) for k, v in host.items()}
loop = pipe(**devin, generator=torch.Generator(device=dev).manual_seed(1), prepare_only=True, num_inference_steps=30,
            guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0 if wl == "config3" else 1.0,
            use_cuda_graph=False)
loop.step(0); loop.step(1)
ops.PROFILE = []
loop.step(2)
torch.cuda.synchronize()
rows = {}
for name, work, a, b in ops.PROFILE:
    key = (name,) + tuple(sorted((k, v) for k, v in work.items() if k not in ("flops", "bytes")))
    d = rows.setdefault(key, dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
    d["n"] += 1; d["ms"] += a.elapsed_time(b); d["flops"] += work.get("flops", 0); d["bytes"] += work.get("bytes", 0)
tot = sum(d["ms"] for d in rows.values())
print(f"total {tot:.2f} ms over {sum(d['n'] for d in rows.values())} launches")
for key, d in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    rate = f"{d['flops'] / d['ms'] / 1e9:8.1f} TF/s" if d["flops"] else (f"{d['bytes'] / d['ms'] / 1e6:8.1f} GB/s" if d["bytes"] else "")
    print(f"{d['ms']:8.3f} ms {100 * d['ms'] / tot:5.1f}%  n={d['n']:4d}  {d['ms'] / d['n'] * 1e3:8.1f} us/launch  {rate}  {key}")
