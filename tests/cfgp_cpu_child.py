"""One rank of tests/test_pipeline_parallel_cpu.py (2 ranks, gloo, CPU): the product pipeline's multi-GPU host logic — CFG-parallel
(one branch per rank, pair leader's noise broadcast, one all-gather of eps per step) and data-parallel sharding (full-batch noise,
sliced) — driven by the oracle modules, against the single-process product loop and the oracle loop.  The GPU counterpart is
tools/cfgp_check.py (tests/test_multi_gpu.py, needs 2 GPUs)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from _cpu_loop import OracleAgg, OracleUNet, emulate_ops  # noqa: E402
from _util import build_oracle, make_inputs, rel_l2  # noqa: E402
from instantir_b200 import config as pcfg, parallel, pipeline, schedulers  # noqa: E402
from oracle import config as ocfg, pipeline as opipe, schedulers as osched  # noqa: E402

rank, port = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)
torch.set_grad_enabled(False)
torch.set_num_threads(2)
emulate_ops(setattr)
oc = ocfg.tiny()
ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
B = 2
inp = make_inputs(oc, B=B, h=8, w=8)
common = dict(num_inference_steps=3, guidance_scale=7.0, preview_start=0.0)
ref = opipe.restore_latents(
    ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"], prompt_embeds=inp["prompt_embeds"],
    negative_prompt_embeds=inp["negative_prompt_embeds"], pooled_prompt_embeds=inp["pooled_prompt_embeds"],
    negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"], ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"],
    generator=torch.Generator().manual_seed(42), **common)
pipe = pipeline.InstantIRPipeline(OracleUNet(ounet, pcfg.ModelConfig(**oc.to_dict())), OracleAgg(oagg), schedulers.DDPMScheduler())
kw = dict(prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
          pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
          previewer_scheduler=schedulers.LCMSingleStepScheduler(), use_cuda_graph=False, overlap_streams=False, **common)
# ---- CFG-parallel: both ranks hold all B images and run one branch each.  The cond rank arrives with a DIFFERENT generator
# state, then with none at all: the pair leader's draws are broadcast, so the result must equal the run seeded 42
cp = parallel.CFGParallel()
assert cp.branch == rank
rows = []
out = pipe(image=inp["image"], ip_adapter_image_embeds=[inp["ip"]], generator=torch.Generator().manual_seed(42 + 1000 * rank),
           cfg_parallel=cp, save_preview_row=True, return_dict=True, **kw)
e_cfgp = rel_l2(out.images, ref)
n_rows = len(out.preview_rows)
out2 = pipe(image=inp["image"], ip_adapter_image_embeds=[inp["ip"]], generator=torch.Generator().manual_seed(42) if rank == 0 else None,
            cfg_parallel=cp, **kw).images
e_cfgp2 = rel_l2(out2, ref)
# ---- data-parallel: rank r restores image r, drawing the full-batch noise and keeping its slice
sl, _ = parallel.partition(B, 2, rank, cfg_parallel=False)
kw_dp = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in kw.items()}
out_dp = pipe(image=inp["image"][sl], ip_adapter_image_embeds=[inp["ip"][:, sl]], generator=torch.Generator().manual_seed(42),
              dp_shard=(B, sl), **kw_dp).images
e_dp = rel_l2(out_dp, ref[sl])
try:
    pipe(image=inp["image"][sl], ip_adapter_image_embeds=[inp["ip"][:, sl]], generator=None, dp_shard=(B, sl), **kw_dp)
    refused = False
except ValueError:
    refused = True
res = torch.tensor([e_cfgp, e_cfgp2, e_dp])
dist.all_reduce(res, op=dist.ReduceOp.MAX)
# the reference keeps the COND chunk of every preview (pipelines/sdxl_instantir.py:1564-1567): only the cond rank holds rows
print(f"CPU_PARALLEL_CHECK rank={rank} cfgp={float(res[0]):.3e} cfgp_unseeded_partner={float(res[1]):.3e} dp={float(res[2]):.3e} "
      f"preview_rows={n_rows} dp_without_generator_refused={refused}")
assert float(res.max()) < 2e-5 and refused and n_rows == (3 if rank == 1 else 0)
dist.barrier()
dist.destroy_process_group()
