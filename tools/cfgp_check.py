"""Run under torchrun with 2 ranks: CFG-parallel (one branch per rank, one all-gather of eps per step)
must reproduce the single-GPU CFG run; DP sharding must reproduce the matching slice of a batched run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from _util import build_oracle, export_state, make_inputs, rel_l2
from instantir_b200 import config as pcfg, parallel, weights
from instantir_b200.aggregator import Aggregator
from instantir_b200.pipeline import InstantIRPipeline
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
from instantir_b200.unet import UNet2DConditionModel
from oracle import config as ocfg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
torch.set_grad_enabled(False)
oc = ocfg.tiny(); alpha = 8.0
ounet, oagg = build_oracle(oc, 0, alpha)
usd, ulora = export_state(ounet); asd, _ = export_state(oagg)
pc = pcfg.ModelConfig(**oc.to_dict())
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, dev, lora=ulora, lora_scale=alpha / oc.lora_rank), dev, prec)
agg = Aggregator(pc, weights.StateDictSource(asd, dev), dev, prec)
pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
B = 2
inp = make_inputs(oc, B=B, h=32, w=32)
kw = dict(prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
          pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
          num_inference_steps=3, guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0)
full = pipe(image=inp["image"], ip_adapter_image_embeds=[inp["ip"]], generator=torch.Generator().manual_seed(42), **kw).images
# --- CFG-parallel: both ranks hold all B images, each runs one branch
cp = parallel.CFGParallel()
# the cond rank arrives with a DIFFERENT generator state (and, second call, with none at all): the pair leader's
# draws are broadcast, so the result must still equal the single-GPU run seeded 42
out = pipe(image=inp["image"], ip_adapter_image_embeds=[inp["ip"]], generator=torch.Generator().manual_seed(42 + 1000 * rank),
           cfg_parallel=cp, **kw).images
e_cfgp = rel_l2(out, full)
# --- data parallel: rank r restores image r of the batch, drawing the full-batch noise and slicing
sl, _ = parallel.partition(B, world, rank, cfg_parallel=False)
kw_dp = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in kw.items()}
out_dp = pipe(image=inp["image"][sl], ip_adapter_image_embeds=[inp["ip"][:, sl]], generator=torch.Generator().manual_seed(42),
              dp_shard=(B, sl), **kw_dp).images
e_dp = rel_l2(out_dp, full[sl])
res = torch.tensor([e_cfgp, e_dp], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"MULTI_GPU_CHECK precision={prec} cfg_parallel_vs_single={float(res[0]):.3e} dp_shard_vs_batched={float(res[1]):.3e}")
    tol = 1e-5 if prec == "fp32" else 2e-2 if prec == "bf16" else 5e-3  # 16-bit: different batch partition -> different tiles/rounding
    assert float(res[0]) < tol and float(res[1]) < tol
dist.barrier(); dist.destroy_process_group()
