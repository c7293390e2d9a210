"""Model-level parity on the GPU, through the reference-shaped Python API which calls the C ABI.

Tolerances are the north star's: per-step latent relative L2 <= 1e-4 in the fp32 check mode and <= 1e-2 in the
headline 16-bit precision, which is fp16 (the reference's own, infer.py:119).  bf16 misses that bar at CFG 7 and is
tested as out of spec (xfail at 1e-2 + a regression envelope).  Single-module outputs (not yet damped by the
scheduler step) get 3e-3 in fp16 and 2e-2 in bf16.
Golden comparisons use vectors produced by running the reference's own files (tests/golden)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import build_oracle, export_state, make_inputs, rel_l2  # noqa: E402
from seeding import seeded_init  # noqa: E402

from instantir_b200 import config as pcfg  # noqa: E402
from instantir_b200 import weights  # noqa: E402
from instantir_b200.aggregator import Aggregator  # noqa: E402
from instantir_b200.nn import Runtime  # noqa: E402
from instantir_b200.pipeline import InstantIRPipeline  # noqa: E402
from instantir_b200.resampler import Resampler  # noqa: E402
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler  # noqa: E402
from instantir_b200.unet import UNet2DConditionModel  # noqa: E402
from oracle import config as ocfg  # noqa: E402
from oracle import model as om  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
from oracle import schedulers as osched  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
TOL = {"fp32": 1e-4, "bf16": 2e-2, "fp16": 3e-3}
torch.set_grad_enabled(False)


def _pcfg_from(oc):
    return pcfg.ModelConfig(**oc.to_dict())


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_unet_base_mode_vs_reference_min_sdxl_vector(precision):
    """product UNet (no adapter: text-only cross-attention) vs the output of the reference's
    module/min_sdxl.py forward (tests/golden/min_sdxl.pt)."""
    g = torch.load(os.path.join(G, "min_sdxl.pt"))
    oc = ocfg.StepConfig(block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                         num_attention_heads=(1, 2, 4), cross_attention_dim=2048, addition_time_embed_dim=32,
                         pooled_dim=64, time_embed_dim=1280)
    ounet = seeded_init(om.UNet2DConditionModel(oc), g["seeds"]["unet"])
    sd, _ = export_state(ounet)
    unet = UNet2DConditionModel(_pcfg_from(oc), weights.StateDictSource(sd, DEV), DEV, precision, adapter=False)
    out = unet(g["sample"].to(DEV), torch.tensor(g["t"]), g["text"].to(DEV),
               added_cond_kwargs={"text_embeds": g["pooled"].to(DEV), "time_ids": g["time_ids"].to(DEV)})[0]
    torch.cuda.synchronize()
    assert rel_l2(out, g["unet_out"]) < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_aggregator_vs_reference_forward_vector(precision):
    """product Aggregator vs the output of the reference's module/aggregator.py forward."""
    g = torch.load(os.path.join(G, "aggregator.pt"))
    oc = ocfg.StepConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["cfg"].items()})
    oagg = om.Aggregator(oc)
    om.remove_attn2(oagg)
    seeded_init(oagg, g["seed"])
    sd, _ = export_state(oagg)
    agg = Aggregator(_pcfg_from(oc), weights.StateDictSource(sd, DEV), DEV, precision)
    i = g["inputs"]
    down, mid = agg(i["sample"].to(DEV), torch.tensor(i["t"]), None, controlnet_cond=i["cond"].to(DEV),
                    added_cond_kwargs={"text_embeds": i["pooled"].to(DEV), "time_ids": i["time_ids"].to(DEV)})
    torch.cuda.synchronize()
    assert len(down) == 9
    for a, b in zip(down, g["down"]):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_l2(a, b) < TOL[precision]
    assert rel_l2(mid, g["mid"]) < TOL[precision]


NOT_YET_RUN_ON_GPU = ("written after the round's GPU budget was spent: the code under test is exercised on the CPU only "
                      "(tests/test_host_logic.py covers its weight source and validation, tests/test_model_host_cpu.py runs the same "
                      "scenario with emulated kernels); non-strict, so an XPASS in the driver's log is the GPU verification")


@pytest.mark.xfail(reason=NOT_YET_RUN_ON_GPU, strict=False)
def test_aggregator_from_unet_then_load_state_dict_fp32():
    """pipelines/sdxl_instantir.py:320-322 + infer.py:142-144 on the product objects: Aggregator.from_unet(unet) gives
    exactly-zero residuals (zero 1x1 heads, module/aggregator.py:980-983) and conditioning_scale only scales them;
    load_state_dict(aggregator.pt keys) then makes it identical to an Aggregator constructed from that state dict."""
    g = torch.load(os.path.join(G, "aggregator.pt"))
    oc = ocfg.StepConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["cfg"].items()})
    ounet = seeded_init(om.load_adapter(om.UNet2DConditionModel(oc)), 11)
    usd, _ = export_state(ounet)
    pc = _pcfg_from(oc)
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV), DEV, "fp32")
    agg = Aggregator.from_unet(unet)
    i = g["inputs"]
    kw = dict(controlnet_cond=i["cond"].to(DEV), added_cond_kwargs={"text_embeds": i["pooled"].to(DEV), "time_ids": i["time_ids"].to(DEV)})
    down, mid = agg(i["sample"].to(DEV), torch.tensor(i["t"]), None, **kw)
    torch.cuda.synchronize()
    assert len(down) == 9 and not any(bool(d.any()) for d in down) and not bool(mid.any())
    oagg = om.Aggregator(oc)
    om.remove_attn2(oagg)
    seeded_init(oagg, g["seed"])
    sd, _ = export_state(oagg)
    res = agg.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys and agg.weights_version == 1
    down, mid = agg(i["sample"].to(DEV), torch.tensor(i["t"]), None, **kw)
    ref = Aggregator(pc, weights.StateDictSource(sd, DEV), DEV, "fp32")
    rdown, rmid = ref(i["sample"].to(DEV), torch.tensor(i["t"]), None, **kw)
    half, _ = ref(i["sample"].to(DEV), torch.tensor(i["t"]), None, conditioning_scale=0.5, **kw)
    torch.cuda.synchronize()
    for a, b, c, d in zip(down, rdown, g["down"], half):
        assert torch.equal(a, b) and rel_l2(a, c) < 1e-4 and rel_l2(d, 0.5 * c) < 1e-4
    assert torch.equal(mid, rmid) and rel_l2(mid, g["mid"]) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_resampler_vs_reference_vector(precision):
    g = torch.load(os.path.join(G, "resampler.pt"))
    ors = seeded_init(om.Resampler(dim=128, depth=2, dim_head=64, heads=2, num_queries=16, embedding_dim=64,
                                   output_dim=256, ff_mult=4), g["seed"])
    cfg = pcfg.tiny()
    sd = {"rs." + k: v for k, v in ors.state_dict().items()}
    rs = Resampler(Runtime(DEV, precision), weights.StateDictSource(sd, DEV), "rs", cfg)
    x = g["x"].to(DEV)
    out = rs(x.reshape(x.shape[0] * x.shape[1], *x.shape[2:]))
    torch.cuda.synchronize()
    assert rel_l2(out, g["out"]) < (1e-4 if precision == "fp32" else 1e-2)


def _run_pair(precision, cfg_name="tiny", steps=2, B=1, h=32, preview_start=0.0, cge=1.0, graph=True, guidance=7.0,
              timesteps=None, adastep=False, **pipe_kw):
    oc = getattr(ocfg, cfg_name)()
    alpha = 8.0
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=alpha)
    inp = make_inputs(oc, B=B, h=h, w=h)
    rec_o = {}
    ref = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], num_inference_steps=None if timesteps else steps,
        timesteps=timesteps, guidance_scale=guidance,
        preview_start=preview_start, control_guidance_end=cge, generator=torch.Generator().manual_seed(42), record=rec_o,
        adastep_restore=adastep)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = _pcfg_from(oc)
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=alpha / oc.lora_rank), DEV, precision)
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, precision)
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
    pipe.prepare_previewers()
    rec_p = {}
    out = pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
               ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=None if timesteps else steps, timesteps=timesteps,
               guidance_scale=guidance,
               previewer_scheduler=LCMSingleStepScheduler(), preview_start=preview_start, control_guidance_end=cge,
               generator=torch.Generator().manual_seed(42), use_cuda_graph=graph, record=rec_p, adastep_restore=adastep, **pipe_kw)
    torch.cuda.synchronize()
    return ref, rec_o, out.images, rec_p


@pytest.mark.parametrize("precision,graph", [("fp32", False), ("fp32", True)])
def test_full_step_config1_fp32_check_mode(precision, graph):
    """BASELINE config 1: scaled-down UNet + aggregator + IP-adapter + LoRA previewer, 256², 2 steps,
    CFG 7 — per-step latent relative L2 vs the oracle <= 1e-4 in the fp32 check mode."""
    ref, rec_o, out, rec_p = _run_pair(precision, graph=graph)
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-4, f"step {i}"
    for a, b in zip(rec_p["preview"], rec_o["preview"]):
        assert (b is None) == (a is None)
        if b is not None:
            assert rel_l2(a, b) < 1e-4
    assert rel_l2(out, ref) < 1e-4


BF16_OUT_OF_SPEC = ("bf16 operands are OUT OF SPEC for the north star's <= 1e-2 per-step latent bar at CFG 7 (measured 1.4e-2 on the "
                   "30-step spacing, 2.5e-2 on config 1's 2-step schedule): 8-bit-mantissa rounding through ~60 sequential GEMM stages "
                   "gives ~1.2e-2 per branch, which CFG 7 amplifies (DESIGN.md §4).  The headline precision is fp16 (the reference's "
                   "own, infer.py:119), which meets the bar; precision='bf16' remains available and is reported as out of spec.")


@pytest.mark.xfail(reason=BF16_OUT_OF_SPEC, strict=False)
@pytest.mark.parametrize("graph", [False, True])
def test_full_step_bf16_on_30_step_spacing(graph):
    """bf16 tcgen05 path on the first steps of the real 30-step schedule (t = 958, 925 -> prev 925, 892), CFG 7, vs the
    fp32 oracle, held to the north star's 1e-2: an expected failure, kept so the gap stays measured (not hidden behind a
    widened tolerance).  A third timestep is listed only so that 925's predecessor is 892 as in the 30-step run."""
    ref, rec_o, out, rec_p = _run_pair("bf16", graph=graph, timesteps=[958, 925, 892])
    errs = [rel_l2(rec_p["latents"][i], rec_o["latents"][i]) for i in range(2)]
    print("bf16 per-step latent rel L2:", errs)
    assert max(errs) < 1e-2, errs


def test_bf16_stays_within_its_measured_envelope():
    """bf16 is out of spec, but it must not drift further: 2e-2 on the 30-step spacing (regression guard, NOT a parity
    claim — parity is claimed for fp16 and the fp32 check mode only)."""
    ref, rec_o, out, rec_p = _run_pair("bf16", graph=True, timesteps=[958, 925, 892])
    for i in range(2):
        assert rel_l2(rec_p["latents"][i], rec_o["latents"][i]) < 2e-2, f"step {i}"


@pytest.mark.parametrize("graph", [False, True])
def test_full_step_fp16_on_30_step_spacing_meets_north_star(graph):
    """fp16 operands (the reference's own inference precision, infer.py:119) on the tcgen05 path, CFG 7,
    first steps of the 30-step schedule: per-step latent relative L2 vs the fp32 oracle <= 1e-2."""
    ref, rec_o, out, rec_p = _run_pair("fp16", graph=graph, timesteps=[958, 925, 892])
    for i in range(2):
        assert rel_l2(rec_p["latents"][i], rec_o["latents"][i]) < 1e-2, f"step {i}"


def test_full_step_config1_fp16():
    """config 1's own 2-step schedule in fp16: <= 1e-2 even across the 500-timestep jump."""
    ref, rec_o, out, rec_p = _run_pair("fp16", graph=True)
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-2, f"step {i}"


def test_bf16_without_cfg_amplification_meets_1e2():
    """same two steps with guidance_scale = 1 (single branch, no CFG amplification): <= 1e-2."""
    ref, rec_o, out, rec_p = _run_pair("bf16", graph=True, timesteps=[958, 925, 892], guidance=1.0)
    for i in range(2):
        assert rel_l2(rec_p["latents"][i], rec_o["latents"][i]) < 1e-2, f"step {i}"


@pytest.mark.xfail(reason=BF16_OUT_OF_SPEC, strict=False)
def test_full_step_config1_bf16_coarse_schedule():
    """config 1's own 2-step schedule (t = 501, 1) in bf16, held to 1e-2: expected failure (the 500-timestep jump
    multiplies the eps error by d x_prev / d eps = 1.62, against 0.37 on the 30-step spacing)."""
    ref, rec_o, out, rec_p = _run_pair("bf16", graph=True)
    errs = [rel_l2(a, b) for a, b in zip(rec_p["latents"], rec_o["latents"])]
    print("bf16 config-1 per-step latent rel L2:", errs)
    assert max(errs) < 1e-2, errs


def test_batch_of_two_images_fp32():
    """B = 2 images (CFG batch 4): every per-sample structure (temb row-vectors, adaLN, cond_scale,
    per-image GroupNorm, multi-image conv tiles) vs the oracle."""
    ref, rec_o, out, rec_p = _run_pair("fp32", B=2, graph=True)
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-4, f"step {i}"


def test_repeated_calls_reuse_graphs_and_stay_correct_fp32():
    """a second image through the same pipeline object (cached static buffers + CUDA graphs, context
    caches refreshed in place) must be restored as correctly as the first."""
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = _pcfg_from(oc)
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=8.0 / oc.lora_rank), DEV, "fp32")
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "fp32")
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
    for seed, B in ((1234, 1), (99, 1), (7, 2), (1234, 1)):
        inp = make_inputs(oc, B=B, h=32, w=32, seed=seed)
        kw = dict(prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                  pooled_prompt_embeds=inp["pooled_prompt_embeds"],
                  negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"], num_inference_steps=2,
                  guidance_scale=7.0, preview_start=0.0)
        ref = opipe.restore_latents(ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
                                    ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"],
                                    generator=torch.Generator().manual_seed(5), **kw)
        out = pipe(image=inp["image"], ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(),
                   generator=torch.Generator().manual_seed(5), **kw).images
        torch.cuda.synchronize()
        assert rel_l2(out, ref) < 1e-4, (seed, B)


def test_step_shapes_no_preview_and_unet_only_fp32():
    """preview_start=1 (aggregator fed the LQ latent) and control_guidance_end=0.5 (second step UNet-only)."""
    ref, rec_o, out, rec_p = _run_pair("fp32", preview_start=1.0, cge=0.5, graph=True)
    for a, b in zip(rec_p["latents"], rec_o["latents"]):
        assert rel_l2(a, b) < 1e-4


@pytest.mark.parametrize("graph", [False, True])
def test_aggregator_one_step_ahead_fp32(graph):
    """No previewer: the Aggregator of step i+1 runs beside the UNet of step i (pipeline.py, `agg_ahead`).  Four
    steps cover the stand-alone first Aggregator, both residual sets, and the last step (nothing to run ahead);
    with control_guidance_end = 0.75 the fourth step is UNet-only.  Against the oracle, and bit-identical to the
    in-order schedule (same kernels on the same inputs, only enqueued earlier)."""
    for cge in (1.0, 0.75):
        ref, rec_o, out, rec_p = _run_pair("fp32", steps=4, preview_start=1.0, cge=cge, graph=graph, agg_ahead=True)
        for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
            assert rel_l2(a, b) < 1e-4, f"cge {cge} step {i}"
        _, _, out_inorder, _ = _run_pair("fp32", steps=4, preview_start=1.0, cge=cge, graph=graph, agg_ahead=False)
        assert torch.equal(out, out_inorder)


def test_aggregator_ahead_mixed_with_preview_steps_bf16():
    """preview_end = 0.5: steps 0-1 preview (Aggregator fed the preview latent, in order), steps 2-3 do not (run
    ahead).  bf16 tcgen05 path with the folded LayerNorm; bit-identical to the in-order schedule."""
    oc = ocfg.tiny()
    alpha = 8.0
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=alpha)
    inp = make_inputs(oc, B=1, h=32, w=32)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = _pcfg_from(oc)
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=alpha / oc.lora_rank), DEV, "bf16")
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "bf16")
    outs = []
    for ahead in (True, False):
        pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
        outs.append(pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                         pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                         ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=4, guidance_scale=7.0,
                         previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0, preview_end=0.5,
                         generator=torch.Generator().manual_seed(42), agg_ahead=ahead).images.clone())
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])


def test_guidance_rescale_fp32():
    """guidance_rescale = 0.7 (rescale_noise_cfg, pipelines/sdxl_instantir.py:181-192): per-sample std rescaling of the
    guided prediction through iir_cfg_rescale, against the oracle's restatement; batch 2 so the two samples get
    different factors."""
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
    inp = make_inputs(oc, B=2, h=32, w=32)
    kw = dict(num_inference_steps=2, guidance_scale=7.0, preview_start=0.0)
    ref = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], generator=torch.Generator().manual_seed(42),
        guidance_rescale=0.7, **kw)
    plain = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], generator=torch.Generator().manual_seed(42), **kw)
    assert rel_l2(ref, plain) > 1e-2  # the option changes the result, so the check below is not vacuous
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = _pcfg_from(oc)
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=8.0 / oc.lora_rank), DEV, "fp32")
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "fp32")
    out = InstantIRPipeline(unet, agg, DDPMScheduler())(
        image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(),
        generator=torch.Generator().manual_seed(42), guidance_rescale=0.7, **kw).images
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-4


ADASTEP_UNVERIFIED = ("row f4 (adastep_restore) is NOT re-verified on a GPU: the only GPU run of this test (round 2, before the commit "
                      "that introduced it) failed on preview_factor at the first previewing step; the host loop was changed afterwards "
                      "(loop.last_previewed in instantir_b200/pipeline.py: pred_x0 is compared with the last preview latent the "
                      "Aggregator was fed, not with the LQ latent) and now reproduces the oracle's bookkeeping on the CPU "
                      "(tests/test_pipeline_loop_cpu.py::test_adastep_restore_bookkeeping_matches_oracle), but the round's GPU budget was "
                      "spent before this test could run again.  Non-strict: an XPASS in the driver's log is the GPU verification.")


@pytest.mark.xfail(reason=ADASTEP_UNVERIFIED, strict=False)
@pytest.mark.parametrize("precision,tol,preview_start", [("fp32", 1e-4, 0.0), ("fp32", 1e-4, 0.5), ("fp16", 1e-2, 0.0)])
def test_adastep_restore(precision, tol, preview_start):
    """adastep_restore (pipelines/sdxl_instantir.py:1636-1644): the per-image preview_factor
    = |preview - pred_x0|^2 / |preview - previous preview|^2 scales (clamped) the next step's residuals; B = 2 so the two
    images get different factors.  preview_start = 0.5: the first two steps feed the Aggregator the LQ latent."""
    ref, rec_o, out, rec_p = _run_pair(precision, steps=4, B=2, preview_start=preview_start, adastep=True, graph=True)
    for i, (a, b) in enumerate(zip(rec_p["preview_factor"], rec_o["preview_factor"])):
        a = a.cpu()
        # a step whose cond_scale fell below the 0.1 gate re-uses the previous preview: |preview - previous preview|^2 = 0 and
        # the factor is +inf in the reference too (then clamped to the step's scale)
        assert torch.equal(torch.isinf(a), torch.isinf(b)), f"preview_factor, step {i}: {a.tolist()} vs {b.tolist()}"
        fin = torch.isfinite(b)
        if bool(fin.any()):
            assert rel_l2(a[fin], b[fin]) < 10 * tol, f"preview_factor, step {i}: {a.tolist()} vs {b.tolist()}"
    assert float((rec_o["preview_factor"][1] - 1.0).abs().min()) > 1e-3  # the factor really moves: the check is not vacuous
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < tol, f"step {i}"


def test_guidance_scale_le_1_disables_cfg_fp32():
    ref, _, out, _ = _run_pair("fp32", steps=1, guidance=1.0, graph=False)
    assert rel_l2(out, ref) < 1e-4


# ------------------------------------------------------------------ SDXL widths (BASELINE configs 2-5)
def _device_seeded_init(module, seed, scale_keys=()):
    """seeding.seeded_init's distribution, drawn on the GPU (4.4 G parameters in seconds instead of minutes on the host)
    and copied into the CPU oracle module; returns the same tensors as a device state dict for the product."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    sd = {}
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        t = torch.randn(p.shape, generator=g, device=DEV, dtype=torch.float32)
        if p.ndim >= 2:
            t.mul_(p[0].numel() ** -0.5)
        else:
            t.mul_(0.05)
            if "norm" in name and name.endswith("weight"):
                t.add_(1.0)
        if any(name.startswith(k[0]) and name.endswith(k[1]) for k in scale_keys):
            t.mul_(0.5)
        p.data.copy_(t)
        sd[name] = t
    return sd


@pytest.fixture(scope="module")
def sdxl_width_oracle_step():
    """ONE UNet + Aggregator step (BASELINE config 2's shape: previewer off, CFG 7, first step of the 30-step schedule,
    t = 958 -> 925) of the CPU oracle at FULL SDXL widths (2.57 G + 1.0 G parameters, IP-adapter processors and
    Resampler installed), latent 32x32 so that the fp32 CPU run takes seconds."""
    oc = ocfg.sdxl()
    with torch.device("meta"):
        ounet = om.load_adapter(om.UNet2DConditionModel(oc))
        oagg = om.Aggregator(oc)
        om.remove_attn2(oagg)
    ounet.to_empty(device="cpu")
    oagg.to_empty(device="cpu")
    sd_u = _device_seeded_init(ounet, 0)
    # keep the injected residuals O(1) relative to the skips they are added to (as _util.build_oracle does)
    sd_a = _device_seeded_init(oagg, 1, scale_keys=[("controlnet_down_blocks.", ".1.weight"), ("controlnet_mid_block.1", "weight")])
    inp = make_inputs(oc, B=1, h=32, w=32)
    rec = {}
    opipe.restore_latents(
        ounet.eval(), oagg.eval(), osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], num_inference_steps=30, guidance_scale=7.0,
        preview_start=1.0, generator=torch.Generator().manual_seed(42), record=rec, max_steps=1)
    del ounet, oagg
    return dict(cfg=oc, inp=inp, sd_u=sd_u, sd_a=sd_a, latents=rec["latents"][0], pred_x0=rec["pred_x0"][0])


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-2), ("fp32", 1e-4)])
def test_sdxl_width_step_vs_oracle(sdxl_width_oracle_step, precision, tol):
    """The product's UNet + Aggregator step at SDXL widths against the CPU ORACLE on the same weights, seeds and
    inputs (not against another mode of the product): per-step latent relative L2 <= 1e-2 in the headline precision
    (fp16 operands on the tcgen05 path) and <= 1e-4 in the fp32 check mode — the north star's bar, at the widths,
    head counts, GEMM / conv shapes and tile choices of BASELINE configs 2-5 (latent 32x32: M = 2048 ... 128 rows)."""
    o = sdxl_width_oracle_step
    pc = _pcfg_from(o["cfg"])
    unet = UNet2DConditionModel(pc, weights.StateDictSource(o["sd_u"], DEV), DEV, precision)
    agg = Aggregator(pc, weights.StateDictSource(o["sd_a"], DEV), DEV, precision)
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler())
    inp, rec = o["inp"], {}
    loop = pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=30, guidance_scale=7.0,
                previewer_scheduler=LCMSingleStepScheduler(), preview_start=1.0, generator=torch.Generator().manual_seed(42),
                record=rec, prepare_only=True)
    loop.step(0)
    torch.cuda.synchronize()
    e_lat, e_x0 = rel_l2(rec["latents"][0], o["latents"]), rel_l2(rec["pred_x0"][0], o["pred_x0"])
    print(f"SDXL-width step vs oracle [{precision}]: latents {e_lat:.3e}, pred_x0 (guided eps) {e_x0:.3e}")
    assert e_lat < tol
    assert e_x0 < (5e-2 if precision == "fp16" else 5e-4)  # x0 = (x - sqrt(1-abar) eps)/sqrt(abar): eps error / 0.087 at t = 958
    del unet, agg, pipe, loop
    torch.cuda.empty_cache()


def _full_models(precision):
    import bench

    cfg = pcfg.sdxl()
    return bench.build_models(cfg, DEV, precision, with_lora=False), cfg


def test_full_size_1024_invariants_and_fp32_check_mode():
    """At BASELINE's full size (SDXL widths, 1024² -> latent 128², CFG batch 2) the CPU oracle needs minutes (the oracle
    comparison at SDXL widths is test_sdxl_width_step_vs_oracle, at latent 32²), so this test adds the size-independent
    properties on identical random-init weights:
      (1) residuals scaled by cond_scale = 0 leave eps bit-identical to the no-residual forward
          (pipelines/sdxl_instantir.py:1602-1603: stale residuals x 0);
      (2) the same launch sequence is bit-deterministic;
      (3) supplementary: the fp16 tcgen05 path agrees with the fp32 check mode (oracle-exact in the test above) on the
          aggregator residuals and on eps of one UNet+aggregator step at the full 128² size."""
    import bench

    cfg = pcfg.sdxl()
    host = bench.host_inputs(cfg, 1, 128)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 128, 128, generator=g).to(DEV)
    t = torch.tensor(501)
    text = torch.cat([host["negative_prompt_embeds"], host["prompt_embeds"]]).to(DEV)
    added = {"text_embeds": torch.cat([host["negative_pooled_prompt_embeds"], host["pooled_prompt_embeds"]]).to(DEV),
             "time_ids": torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]] * 2, device=DEV),
             "image_embeds": [torch.cat([host["ip_adapter_image_embeds"][0], host["ip_adapter_image_embeds"][1]]).unsqueeze(1).to(DEV)]}
    img = torch.cat([host["image"]] * 2).to(DEV)
    res = {}
    for prec in ("fp16", "fp32"):
        (unet, agg), _ = _full_models(prec)
        down, mid = agg(img, t, text, controlnet_cond=x, added_cond_kwargs=added)
        eps = unet(x, t, text, added_cond_kwargs=added, down_block_additional_residuals=down,
                   mid_block_additional_residual=mid)[0]
        if prec == "fp16":
            eps2 = unet(x, t, text, added_cond_kwargs=added, down_block_additional_residuals=down,
                        mid_block_additional_residual=mid)[0]
            assert torch.equal(eps, eps2), "not deterministic"
            zero = torch.zeros(2, device=DEV)
            eps0 = unet(x, t, text, added_cond_kwargs=added, down_block_additional_residuals=down,
                        mid_block_additional_residual=mid, additional_residual_scale=zero)[0]
            plain = unet(x, t, text, added_cond_kwargs=added)[0]
            assert torch.equal(eps0, plain), "cond_scale = 0 must equal the UNet-only forward"
            assert rel_l2(plain, eps) > 1e-3, "residual injection has no effect: the check would be vacuous"
        torch.cuda.synchronize()
        res[prec] = ([d.float().cpu() for d in down], mid.float().cpu(), eps.float().cpu())
        del unet, agg
        torch.cuda.empty_cache()
    assert bool(torch.isfinite(res["fp16"][2]).all())
    errs = [rel_l2(a, b) for a, b in zip(res["fp16"][0], res["fp32"][0])] + [rel_l2(res["fp16"][1], res["fp32"][1])]
    e = rel_l2(res["fp16"][2], res["fp32"][2])
    print(f"full-size eps rel L2 fp16 vs fp32 check mode: {e:.3e}; aggregator residuals max {max(errs):.3e}")
    assert max(errs) < 5e-3, errs
    assert e < 5e-3
