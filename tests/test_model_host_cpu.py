"""CPU test of the product's MODEL host code (instantir_b200/{unet,aggregator,attention_processor,resampler,nn,weights}.py):
the real UNet2DConditionModel / Aggregator objects — weight packing, fused QKV, LoRA-merged second weight set, folded-LayerNorm
chain, time-embedding banks, context caches, residual injection in the concat, the opt-in GroupNorm-statistics plumbing — run
with torch emulations of the kernels' CONTRACTS (tests/_ops_emulation.py) and must reproduce the CPU oracle through the whole
denoising loop.  `fp32` exercises the check-mode call pattern (LayerNorm / SIMT conv geometry), `fp16` the tcgen05 call pattern
(16-bit operands, folded LayerNorm, paired epilogues), `fp16+gn` the opt-in fused GroupNorm path of DESIGN.md §3.6.  The kernels
themselves are checked on the GPU (tests/test_kernels_gpu.py, tests/test_model_parity_gpu.py)."""
import pytest
import torch

import _ops_emulation
from _cpu_loop import emulate_ops
from _util import build_oracle, export_state, make_inputs, rel_l2
from instantir_b200 import config as pcfg
from instantir_b200 import pipeline, schedulers, weights
from oracle import config as ocfg
from oracle import pipeline as opipe
from oracle import schedulers as osched


def _oracle_and_product(monkeypatch, precision, fuse_gn=False, B=1, lat=16, steps=2, overlap=False, **kw):
    from instantir_b200.aggregator import Aggregator
    from instantir_b200.unet import UNet2DConditionModel

    emu = _ops_emulation.install(monkeypatch.setattr, fuse_gn=fuse_gn)
    emulate_ops(monkeypatch.setattr)
    oc = ocfg.tiny()
    alpha = 8.0
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=alpha)
    inp = make_inputs(oc, B=B, h=lat, w=lat)
    rec_o = {}
    ref = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"], prompt_embeds=inp["prompt_embeds"],
        negative_prompt_embeds=inp["negative_prompt_embeds"], pooled_prompt_embeds=inp["pooled_prompt_embeds"],
        negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"], ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"],
        num_inference_steps=steps, guidance_scale=7.0, generator=torch.Generator().manual_seed(42), record=rec_o, **kw)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, "cpu", lora=ulora, lora_scale=alpha / oc.lora_rank), "cpu", precision)
    agg = Aggregator(pc, weights.StateDictSource(asd, "cpu"), "cpu", precision)
    assert unet.rt.device.type == "cpu" and unet.rt.gn_fuse == (fuse_gn and precision != "fp32")
    pipe = pipeline.InstantIRPipeline(unet, agg, schedulers.DDPMScheduler())
    rec_p = {}
    out = pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
               ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=steps, guidance_scale=7.0,
               previewer_scheduler=schedulers.LCMSingleStepScheduler(), generator=torch.Generator().manual_seed(42),
               use_cuda_graph=False, overlap_streams=overlap, record=rec_p, **kw)
    return ref, rec_o, out.images, rec_p, emu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_full_step_host_code_matches_oracle(monkeypatch, precision, tol):
    """BASELINE config 1 at latent 16²: previewer (LoRA weight set) + LCM + Aggregator + UNet + CFG 7 + DDPM, 2 steps"""
    ref, rec_o, out, rec_p, emu = _oracle_and_product(monkeypatch, precision, preview_start=0.0)
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < tol, f"step {i}: {rel_l2(a, b):.3e}"
    assert rel_l2(out, ref) < tol
    kinds = [c[0] for c in emu.calls]
    gemms = [c[1] for c in emu.calls if c[0] == "gemm"]
    if precision == "fp16":   # the tcgen05 call pattern: no LayerNorm launches inside the transformer blocks, folded instead
        assert any(g["ln_in"] for g in gemms) and any(g["ln_out"] for g in gemms) and all(g["tc"] for g in gemms)
        assert "groupnorm_apply_sums" not in kinds and not any(g["gn"] for g in gemms)
    else:                     # the fp32 check mode: SIMT GEMMs, LayerNorm kernels
        assert not any(g["tc"] for g in gemms) and "layernorm" in kinds


def test_fused_groupnorm_path_matches_oracle_and_removes_stat_passes(monkeypatch):
    """the opt-in path through the REAL UNet / Aggregator host code: per-step latents within the fp16 bar, most GroupNorms
    served from epilogue statistics, and the same results as the default path up to 16-bit storage effects"""
    # latent 32²: every level (32², 16², 8²) has >= 32 pixels of one image per conv tile, so only the GroupNorms behind conv_in
    # and behind the up path's concat fall back to the two-kernel form (10 of the UNet's 46, 1 of the Aggregator's 21)
    ref, rec_o, out1, rec1, emu1 = _oracle_and_product(monkeypatch, "fp16", fuse_gn=True, lat=32, steps=1, preview_start=0.0)
    for i, (a, b) in enumerate(zip(rec1["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-2, f"step {i}: {rel_l2(a, b):.3e}"
    k1 = [c[0] for c in emu1.calls]
    n_fused, n_two = k1.count("groupnorm_apply_sums"), k1.count("groupnorm")
    assert (n_fused, n_two) == (2 * 36 + 20, 2 * 10 + 1), (n_fused, n_two)  # previewer UNet + Aggregator + UNet
    assert k1.count("memset_zero") > 0
    ref, rec_o, out0, rec0, emu0 = _oracle_and_product(monkeypatch, "fp16", fuse_gn=False, lat=32, steps=1, preview_start=0.0)
    k0 = [c[0] for c in emu0.calls]
    assert k0.count("groupnorm") == n_fused + n_two and "memset_zero" not in k0
    assert rel_l2(out1, out0) < 2e-3


def test_mixed_schedule_and_batch_host_code_fp32(monkeypatch):
    """two images, 4 steps: LQ-fed Aggregator, previewing, UNet-only tail (cached step-invariant context across steps)"""
    ref, rec_o, out, rec_p, _ = _oracle_and_product(monkeypatch, "fp32", B=2, steps=4, preview_start=0.5, control_guidance_end=0.75)
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-4, f"step {i}: {rel_l2(a, b):.3e}"


def test_aggregator_from_unet_and_load_state_dict_host_code(monkeypatch):
    """pipelines/sdxl_instantir.py:320-322 + infer.py:142-144 on the product objects (host code under test, kernels emulated):
    Aggregator.from_unet(unet) gives exactly-zero residuals; load_state_dict(aggregator.pt keys) makes it identical to an
    Aggregator built from that state dict and equal to the oracle's forward; conditioning_scale scales the residuals."""
    from instantir_b200.aggregator import Aggregator
    from instantir_b200.unet import UNet2DConditionModel
    from oracle import model as om

    _ops_emulation.install(monkeypatch.setattr)
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0)
    usd, _ = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, "cpu"), "cpu", "fp32")
    agg = Aggregator.from_unet(unet)
    inp = make_inputs(oc, B=2, h=16, w=16)
    t = torch.tensor(501)
    cond = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    kw = dict(controlnet_cond=cond, added_cond_kwargs={"text_embeds": inp["pooled_prompt_embeds"], "time_ids": inp["time_ids"]})
    down, mid = agg(inp["image"], t, None, **kw)
    assert len(down) == 9 and not any(bool(d.any()) for d in down) and not bool(mid.any())
    res = agg.load_state_dict(asd)
    assert not res.missing_keys and not res.unexpected_keys and agg.weights_version == 1
    down, mid = agg(inp["image"], t, None, **kw)
    ref_agg = Aggregator(pc, weights.StateDictSource(asd, "cpu"), "cpu", "fp32")
    rdown, rmid = ref_agg(inp["image"], t, None, **kw)
    odown, omid = oagg(inp["image"], t, encoder_hidden_states=None, **kw, return_dict=False)
    half, hmid = ref_agg(inp["image"], t, None, conditioning_scale=0.5, **kw)
    for a, b, c, d in zip(down, rdown, odown, half):
        assert torch.equal(a, b) and rel_l2(a, c) < 1e-4 and rel_l2(d, 0.5 * c) < 1e-4
    assert torch.equal(mid, rmid) and rel_l2(mid, omid) < 1e-4 and rel_l2(hmid, 0.5 * omid) < 1e-4
    # strict=False keeps the tensors that are absent from the new state dict
    part = {k: v for k, v in asd.items() if not k.startswith("controlnet_mid_block")}
    res = agg.load_state_dict(part, strict=False)
    assert res.missing_keys and all(k.startswith("controlnet_mid_block") for k in res.missing_keys)
    _, mid2 = agg(inp["image"], t, None, **kw)
    assert torch.equal(mid2, mid)


class _Ev:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass


@pytest.mark.parametrize("fuse_gn", [False, True])
def test_forked_step_host_code_matches_oracle(monkeypatch, fuse_gn):
    """the DEFAULT step schedule (overlap_streams=True: UNet down + mid on a second stream beside the Aggregator, SFT heads on
    that stream too, then the up path) — with streams stubbed to run in program order, the host code must still reproduce
    the oracle; also with the opt-in fused GroupNorm (separate accumulator arenas per model)"""
    import contextlib

    from _cpu_loop import NullStream

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: NullStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Event", _Ev)
    ref, rec_o, out, rec_p, emu = _oracle_and_product(monkeypatch, "fp16", fuse_gn=fuse_gn, steps=2, overlap=True, preview_start=0.5)
    assert len(rec_p["latents"]) == 2
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-2, f"step {i}: {rel_l2(a, b):.3e}"


def test_reference_latents_and_denoising_end_through_the_real_host_code(monkeypatch):
    """the two options through the real UNet / Aggregator objects (emulated kernels): a third static conditioning buffer and
    its own graph slots, a shortened run"""
    refl = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(11))
    ref, rec_o, out, rec_p, _ = _oracle_and_product(monkeypatch, "fp32", steps=4, preview_start=0.5, reference_latents=refl,
                                                     denoising_end=0.75)
    assert len(rec_p["latents"]) == len(rec_o["latents"]) == 3
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-4, f"step {i}: {rel_l2(a, b):.3e}"

