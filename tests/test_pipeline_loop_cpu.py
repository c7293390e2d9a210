"""CPU test of the HOST logic of instantir_b200.pipeline.InstantIRPipeline.__call__ — step masks, the cond_scale gate, which
latent feeds the Aggregator, noise drawing order, CFG / DDPM / LCM coefficient plumbing, adastep_restore bookkeeping — with the
device layer replaced: the UNet and the Aggregator are the ORACLE modules behind the product objects' call surface, and the few
ops the loop and the schedulers launch themselves are torch emulations of the kernels' contracts.  The product loop must then
reproduce the oracle loop (oracle/pipeline.py, the restatement of pipelines/sdxl_instantir.py:1497-1666) step for step.
What this does NOT cover is the kernels (GPU tests) — it pins the orchestration, on the box that has no GPU."""
import pytest
import torch

from _cpu_loop import OracleAgg, OracleUNet, emulate_ops
from _util import build_oracle, make_inputs, rel_l2
from instantir_b200 import config as pcfg
from instantir_b200 import pipeline, schedulers
from oracle import config as ocfg
from oracle import pipeline as opipe
from oracle import schedulers as osched


def _run(monkeypatch, *, B=1, steps=3, guidance=7.0, lat=8, **kw):
    emulate_ops(monkeypatch.setattr)
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
    inp = make_inputs(oc, B=B, h=lat, w=lat)
    okw = dict(kw)
    rec_o = {}
    ref = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"], prompt_embeds=inp["prompt_embeds"],
        negative_prompt_embeds=inp["negative_prompt_embeds"], pooled_prompt_embeds=inp["pooled_prompt_embeds"],
        negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"], ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"],
        num_inference_steps=steps, guidance_scale=guidance, generator=torch.Generator().manual_seed(42), record=rec_o, **okw)
    pc = pcfg.ModelConfig(**oc.to_dict())
    pipe = pipeline.InstantIRPipeline(OracleUNet(ounet, pc), OracleAgg(oagg), schedulers.DDPMScheduler())
    rec_p = {}
    out = pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
               ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=steps, guidance_scale=guidance,
               previewer_scheduler=schedulers.LCMSingleStepScheduler(), generator=torch.Generator().manual_seed(42),
               use_cuda_graph=False, overlap_streams=False, record=rec_p, **kw)
    return ref, rec_o, out, rec_p


def _same_steps(rec_p, rec_o, tol=2e-5):
    assert len(rec_p["latents"]) == len(rec_o["latents"])
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < tol, f"latents, step {i}: {rel_l2(a, b):.3e}"
    for i, (a, b) in enumerate(zip(rec_p["pred_x0"], rec_o["pred_x0"])):
        assert rel_l2(a, b) < tol, f"pred_x0, step {i}"
    for i, (a, b) in enumerate(zip(rec_p["preview"], rec_o["preview"])):
        # the product records a preview only on previewing steps; the oracle's record holds whatever fed the Aggregator (the LQ
        # latent on non-previewing controlled steps, None when the gate was closed)
        if a is not None:
            assert b is not None and rel_l2(a, b) < tol, f"preview, step {i}"


def test_loop_with_previewer_matches_oracle(monkeypatch):
    ref, rec_o, out, rec_p = _run(monkeypatch, steps=3, preview_start=0.0)
    _same_steps(rec_p, rec_o)
    assert rel_l2(out.images, ref) < 2e-5


def test_loop_mixed_step_shapes_matches_oracle(monkeypatch):
    """4 steps: LQ-fed Aggregator (0, 1), previewing (2), UNet-only after control_guidance_end (3); batch of two images"""
    ref, rec_o, out, rec_p = _run(monkeypatch, B=2, steps=4, preview_start=0.5, control_guidance_end=0.75)
    _same_steps(rec_p, rec_o)
    assert [p is not None for p in rec_p["preview"]] == [False, False, True, False]
    assert rel_l2(out.images, ref) < 2e-5


def test_loop_without_cfg_and_with_guidance_rescale_matches_oracle(monkeypatch):
    ref, rec_o, out, rec_p = _run(monkeypatch, steps=2, guidance=1.0, preview_start=0.0)
    _same_steps(rec_p, rec_o)
    ref, rec_o, out, rec_p = _run(monkeypatch, B=2, steps=2, preview_start=0.0, guidance_rescale=0.7)
    _same_steps(rec_p, rec_o)
    assert rel_l2(out.images, ref) < 2e-5


@pytest.mark.parametrize("preview_start", [0.0, 0.5])
def test_adastep_restore_bookkeeping_matches_oracle(monkeypatch, preview_start):
    """adastep_restore (pipelines/sdxl_instantir.py:1636-1644, :1538-1542): the same scenario as the GPU test
    test_model_parity_gpu.py::test_adastep_restore, with the host loop alone under test"""
    ref, rec_o, out, rec_p = _run(monkeypatch, B=2, steps=4, preview_start=preview_start, adastep_restore=True)
    for i, (a, b) in enumerate(zip(rec_p["preview_factor"], rec_o["preview_factor"])):
        assert torch.equal(torch.isinf(a), torch.isinf(b)), f"preview_factor, step {i}: {a.tolist()} vs {b.tolist()}"
        fin = torch.isfinite(b)
        if bool(fin.any()):
            assert rel_l2(a[fin], b[fin]) < 1e-4, f"preview_factor, step {i}: {a.tolist()} vs {b.tolist()}"
    assert float((rec_o["preview_factor"][1] - 1.0).abs().min()) > 1e-3  # the factor really moves
    for i, (a, b) in enumerate(zip(rec_p["latents"], rec_o["latents"])):
        assert rel_l2(a, b) < 1e-4, f"latents, step {i}"
    assert rel_l2(out.images, ref) < 1e-4


def test_denoising_end_and_reference_latents_match_oracle(monkeypatch):
    """denoising_end (pipelines/sdxl_instantir.py:1469-1484): only the timesteps at or above the cut-off run, while the step
    masks keep the full schedule; reference_latents (:1579-1580): controlled steps that do not preview feed the Aggregator this
    latent instead of the LQ one — also as the `preview` of adastep_restore"""
    ref, rec_o, out, rec_p = _run(monkeypatch, steps=4, preview_start=0.0, denoising_end=0.5)
    assert len(rec_o["latents"]) == len(rec_p["latents"]) == 2   # t = 751, 501 of [751, 501, 251, 1]; cut-off 500
    _same_steps(rec_p, rec_o)
    assert rel_l2(out.images, ref) < 2e-5
    refl = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(11))
    ref, rec_o, out, rec_p = _run(monkeypatch, B=2, steps=4, preview_start=0.5, reference_latents=refl)
    _same_steps(rec_p, rec_o)
    ref2, rec_o2, _, _ = _run(monkeypatch, B=2, steps=4, preview_start=0.5)
    assert rel_l2(rec_o["latents"][0], rec_o2["latents"][0]) > 1e-3  # the option changes the result: the check is not vacuous
    ref, rec_o, out, rec_p = _run(monkeypatch, B=2, steps=3, preview_start=1.0, reference_latents=refl, adastep_restore=True)
    for a, b in zip(rec_p["preview_factor"], rec_o["preview_factor"]):
        assert torch.equal(torch.isinf(a), torch.isinf(b)) and rel_l2(a[torch.isfinite(b)], b[torch.isfinite(b)]) < 1e-4
    _same_steps(rec_p, rec_o, tol=1e-4)
