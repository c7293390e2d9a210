"""One eager config-2 denoising step for ncu: prints the number of library launches before the last step
so that `ncu -s <skip> -c <count>` captures exactly one step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from instantir_b200 import _lib, config as pcfg
from instantir_b200.pipeline import InstantIRPipeline
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
torch.set_grad_enabled(False)
dev = "cuda:0"
cfg = pcfg.sdxl()
unet, agg = bench.build_models(cfg, dev, "bf16", with_lora=False)
pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
devin = {k: v.to(dev) for k, v in bench.host_inputs(cfg, 1, 128).items()}
loop = pipe(**devin, generator=torch.Generator(device=dev).manual_seed(1), prepare_only=True, num_inference_steps=30,
            guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=1.0, use_cuda_graph=False)
loop.step(0)
torch.cuda.synchronize()
n0 = _lib.launch_count()
loop.step(1)
torch.cuda.synchronize()
print(f"NSKIP={n0} NCOUNT={_lib.launch_count() - n0}", flush=True)
