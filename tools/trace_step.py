"""In-graph per-kernel breakdown of config-2 steps: torch.profiler (CUPTI) around CUDA-graph replays, so the
durations are the ones the bench sees (warm L2, PDL overlap, no per-launch host gaps).  Not a bench number."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from instantir_b200 import config as pcfg
from instantir_b200.pipeline import InstantIRPipeline
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
torch.set_grad_enabled(False)
dev = "cuda:0"
wl = sys.argv[1] if len(sys.argv) > 1 else "config2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
NSTEP = 4
cfg = pcfg.sdxl()
unet, agg = bench.build_models(cfg, dev, "bf16", with_lora=(wl == "config3"))
pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
devin = {k: v.to(dev) for k, v in bench.host_inputs(cfg, B, 128).items()}
loop = pipe(**devin, generator=torch.Generator(device=dev).manual_seed(1), prepare_only=True, num_inference_steps=30,
            guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0 if wl == "config3" else 1.0)
for i in range(4):
    loop.step(i)
torch.cuda.synchronize()
# the launch sequence of a step is deterministic: an eager step with the profiling hook gives the shape key of
# every GEMM / attention launch in order, so in-graph durations can be attributed per shape
from instantir_b200 import ops
loop_e = pipe(**devin, generator=torch.Generator(device=dev).manual_seed(1), prepare_only=True, num_inference_steps=30,
              guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0 if wl == "config3" else 1.0,
              use_cuda_graph=False)
loop_e.step(0)
ops.PROFILE = []
loop_e.step(1)
torch.cuda.synchronize()
seq_gemm = [(w["key"], w["flops"]) for n, w, _, _ in ops.PROFILE if n in ("gemm_tc", "conv3x3_tc")]
seq_attn = [((w["n_q"], w["n_kv"]), w["flops"]) for n, w, _, _ in ops.PROFILE if n == "attn_tc"]
ops.PROFILE = None
del loop_e
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(NSTEP):
    loop.step(4 + i)
e1.record()
torch.cuda.synchronize()
print(f"untraced: {e0.elapsed_time(e1) / NSTEP:.2f} ms/step")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(NSTEP):
        loop.step(8 + i)
    torch.cuda.synchronize()
rows = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
per_gemm, per_attn = [], []
for ev in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start):
    if "gemm_tc_kernel" in ev.name:
        per_gemm.append(ev.device_time)
    elif "attn_t" in ev.name:
        per_attn.append(ev.device_time)
    name = ev.name.replace("void ", "").replace("iir::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
    k = name.split("(")[0]
    rows[k][0] += 1
    rows[k][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    s, e = ev.time_range.start, ev.time_range.end
    t_min = s if t_min is None else min(t_min, s)
    t_max = e if t_max is None else max(t_max, e)
tot = sum(v[1] for v in rows.values())
print(f"traced span {(t_max - t_min) / NSTEP / 1e3:.2f} ms/step; sum of kernel durations {tot / NSTEP / 1e3:.2f} ms/step")
for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / NSTEP / 1e3:8.3f} ms/step {100 * us / tot:5.1f}%  n/step={n / NSTEP:7.1f}  avg {us / n:8.2f} us  {k[:110]}")

for title, seq, per in (("GEMM / conv", seq_gemm, per_gemm), ("attention", seq_attn, per_attn)):
    if len(per) != NSTEP * len(seq):
        print(f"{title}: cannot attribute per shape ({len(per)} traced launches vs {NSTEP} x {len(seq)})")
        continue
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, us in enumerate(per):
        k, fl = seq[i % len(seq)]
        agg[k][0] += 1; agg[k][1] += us; agg[k][2] += fl
    print(f"--- {title} per shape (in-graph): ms/step, launches/step, us/launch, TFLOP/s")
    for k, (n, us, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{us / NSTEP / 1e3:7.3f} ms  n={n // NSTEP:4d}  {us / n:7.1f} us  {fl / us / 1e6:7.0f} TF/s  {k}")
