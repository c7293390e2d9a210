"""CPU test of the HOST logic of instantir_b200.pipeline.InstantIRPipeline.__call__ — step masks, the cond_scale gate, which
latent feeds the Aggregator, noise drawing order, CFG / DDPM / LCM coefficient plumbing, adastep_restore bookkeeping — with the
device layer replaced: the UNet and the Aggregator are the ORACLE modules behind the product objects' call surface, and the few
ops the loop and the schedulers launch themselves are torch emulations of the kernels' contracts.  The product loop must then
reproduce the oracle loop (oracle/pipeline.py, the restatement of pipelines/sdxl_instantir.py:1497-1666) step for step.
What this does NOT cover is the kernels (GPU tests) — it pins the orchestration, on the box that has no GPU."""
from types import SimpleNamespace

import pytest
import torch

from _util import build_oracle, make_inputs, rel_l2
from instantir_b200 import config as pcfg
from instantir_b200 import ops, pipeline, schedulers
from oracle import config as ocfg
from oracle import pipeline as opipe
from oracle import schedulers as osched


class _UNet:
    """oracle UNet behind the product UNet's call surface (unet.py: forward(..., additional_residual_scale=))"""

    def __init__(self, o, cfg):
        self.o, self.cfg, self.rt = o, cfg, SimpleNamespace(device=torch.device("cpu"))

    def enable_adapters(self):
        self.o.enable_adapters()

    def disable_adapters(self):
        self.o.disable_adapters()

    def refresh_context(self, *a):
        pass

    def __call__(self, x, t_dev, encoder_hidden_states=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, additional_residual_scale=None, return_dict=False):
        t = torch.tensor(int(t_dev.reshape(-1)[0]), dtype=torch.int64)
        emb = self.o.time_embedding(self.o.get_time_embed(sample=x, timestep=t))
        emb = emb + self.o.get_aug_embed(emb=emb, encoder_hidden_states=encoder_hidden_states, added_cond_kwargs=added_cond_kwargs)
        kw = {}
        if down_block_additional_residuals is not None:
            s = additional_residual_scale.view(-1, 1, 1, 1)
            kw = dict(down_block_additional_residuals=[d * s for d in down_block_additional_residuals],
                      mid_block_additional_residual=mid_block_additional_residual * s)
        return self.o(x, t, encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs={"temb": emb},
                      added_cond_kwargs=added_cond_kwargs, return_dict=False, **kw)


class _Agg:
    weights_version = 0

    def __init__(self, o):
        self.o = o

    def __call__(self, sample, t_dev, **kw):
        t = torch.tensor(int(t_dev.reshape(-1)[0]), dtype=torch.int64)
        return self.o(sample, t, **kw)


def _emulate_ops(monkeypatch):
    def step_prologue(latents, x_in, n_rep, *, t, t_dev, cond_scale, cond_scale_dev):
        x_in.copy_(torch.cat([latents] * n_rep, 0))
        t_dev.fill_(t)
        if cond_scale_dev is not None:
            cond_scale_dev.fill_(cond_scale)
        return x_in

    def lcm_step(eps, x, out, *, alpha_prod_t, c_skip, c_out):
        a = torch.tensor(alpha_prod_t, dtype=torch.float32)
        x0 = (x - torch.sqrt(1 - a) * eps.float()) / torch.sqrt(a)
        return out.copy_(c_out * x0 + c_skip * x)

    def cfg_ddpm_step(eu, ec, x, noise, prev, pred_x0, *, guidance, alpha_prod_t, c_x0, c_xt, sigma):
        e = eu.float() if ec is None else eu.float() + guidance * (ec.float() - eu.float())
        a = torch.tensor(alpha_prod_t, dtype=torch.float32)
        x0 = (x - torch.sqrt(1 - a) * e) / torch.sqrt(a)
        pred_x0.copy_(x0)
        return prev.copy_(c_x0 * x0 + c_xt * x + (sigma * noise if noise is not None else 0.0))

    def add_noise(x0, noise, out, *, alpha_prod_t):
        a = torch.tensor(alpha_prod_t, dtype=torch.float32)
        return out.copy_(torch.sqrt(a) * x0 + torch.sqrt(1 - a) * noise)

    def adastep_update(preview, pred_x0, previewer_mean, preview_factor, cond_scale, *, n_rep, next_scale, next_keep):
        # the contract of iir_adastep_update (csrc/sched.cu)
        B = pred_x0.shape[0]
        s0 = (preview.double() - pred_x0.double()).pow(2).flatten(1).sum(1)
        s1 = (preview.double() - previewer_mean.double()).pow(2).flatten(1).sum(1)
        f = s0.float() / s1.float()
        preview_factor.copy_(f)
        previewer_mean.copy_(preview)
        cs = torch.clamp(f, min=0.0).clamp(max=next_scale) * next_keep
        cond_scale.copy_(cs.repeat(n_rep))
        assert cond_scale.numel() == n_rep * B

    def cfg_rescale(eu, ec, out, *, guidance, rescale):
        return out.copy_(opipe.rescale_noise_cfg(eu + guidance * (ec - eu), ec, rescale))

    for name, fn in dict(step_prologue=step_prologue, lcm_step=lcm_step, cfg_ddpm_step=cfg_ddpm_step, add_noise=add_noise,
                         adastep_update=adastep_update, cfg_rescale=cfg_rescale).items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: SimpleNamespace())


def _run(monkeypatch, *, B=1, steps=3, guidance=7.0, lat=16, **kw):
    _emulate_ops(monkeypatch)
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
    pipe = pipeline.InstantIRPipeline(_UNet(ounet, pc), _Agg(oagg), schedulers.DDPMScheduler())
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
