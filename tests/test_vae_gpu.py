"""SURVEY §8 row f1 (first "next" row): VAE decode on the sm_100a kernels against the oracle restatement and the
reference-run Decoder vector; decoded-image PSNR of the whole config-1 path against the oracle (north star: >= 40 dB)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from _util import build_oracle, export_state, make_inputs, rel_l2  # noqa: E402
from seeding import seeded_init  # noqa: E402

from instantir_b200 import config as pcfg  # noqa: E402
from instantir_b200 import weights  # noqa: E402
from instantir_b200.aggregator import Aggregator  # noqa: E402
from instantir_b200.pipeline import InstantIRPipeline  # noqa: E402
from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler  # noqa: E402
from instantir_b200.unet import UNet2DConditionModel  # noqa: E402
from instantir_b200.vae import AutoencoderKL, VaeConfig, postprocess, vae_decoder_param_shapes  # noqa: E402
from oracle import config as ocfg  # noqa: E402
from oracle import pipeline as opipe  # noqa: E402
from oracle import schedulers as osched  # noqa: E402
from oracle import vae as ov  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"
G = os.path.join(HERE, "golden")


def _oracle_vae(cfg, seed=71):
    vae = ov.AutoencoderKLDecoder(cfg)
    seeded_init(vae, seed)
    return vae.eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2), ("fp16", 4e-3)])
def test_decoder_vs_reference_run_vector(precision, tol):
    """product Decoder vs the vector produced by the reference's own Decoder.forward (tests/golden/make_golden_vae.py)"""
    g = torch.load(os.path.join(G, "vae_decoder.pt"))
    cfg = ov.VaeConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["cfg"].items()})
    odec = ov.Decoder(cfg)
    seeded_init(odec, g["seed"])
    sd = {"decoder." + k: v for k, v in odec.state_dict().items()}
    L = cfg.latent_channels
    sd["post_quant_conv.weight"] = torch.eye(L).reshape(L, L, 1, 1)
    sd["post_quant_conv.bias"] = torch.zeros(L)
    vae = AutoencoderKL(VaeConfig(**cfg.to_dict()), weights.StateDictSource(sd, DEV), DEV, precision)
    out = vae.decode(g["z"].to(DEV)).sample
    torch.cuda.synchronize()
    assert out.shape == g["out"].shape and torch.isfinite(out).all()
    assert rel_l2(out, g["out"]) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_autoencoder_decode_vs_oracle(precision, tol):
    """post_quant_conv + decoder, batch 3, non-square latent, weights shared through the state dict"""
    cfg = ov.tiny_vae()
    ovae = _oracle_vae(cfg)
    z = torch.randn(3, 4, 16, 24, generator=torch.Generator().manual_seed(5))
    ref = ovae.decode(z)
    vae = AutoencoderKL(VaeConfig(**cfg.to_dict()), weights.StateDictSource(ovae.state_dict(), DEV), DEV, precision)
    assert set(vae_decoder_param_shapes(vae.config)) == set(ovae.state_dict())
    assert all(tuple(v.shape) == tuple(vae_decoder_param_shapes(vae.config)[k]) for k, v in ovae.state_dict().items())
    out = vae.decode(z.to(DEV), return_dict=False)[0]
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < tol
    with pytest.raises(ValueError):
        vae.decode(torch.zeros(1, 3, 8, 8, device=DEV))


_BF16_XFAIL = pytest.mark.xfail(strict=False, reason="bf16 is out of spec for the north star's parity bars at CFG 7 (DESIGN.md §4); "
                                "the headline precision is fp16")


@pytest.mark.parametrize("precision,min_psnr,preview_start", [
    ("fp32", 60.0, 0.0), ("fp16", 40.0, 0.0), ("fp16", 40.0, 1.0),
    pytest.param("bf16", 40.0, 0.0, marks=_BF16_XFAIL), ("bf16", 40.0, 1.0)])
def test_pipeline_decoded_image_psnr_vs_oracle(precision, min_psnr, preview_start):
    """BASELINE config 1 end to end: denoising loop + VAE decode through the public API (output_type='pt') against the
    oracle loop + oracle VAE on the same weights and seeds.  North star: decoded-image PSNR >= 40 dB."""
    oc = ocfg.tiny()
    alpha = 8.0
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=alpha)
    vcfg = ov.tiny_vae()
    ovae = _oracle_vae(vcfg)
    inp = make_inputs(oc, B=1, h=32, w=32)
    kw = dict(num_inference_steps=2, guidance_scale=7.0, preview_start=preview_start)
    ref_lat = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], generator=torch.Generator().manual_seed(42), **kw)
    ref_img = ov.latents_to_image(ovae, ref_lat)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=alpha / oc.lora_rank), DEV, precision)
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, precision)
    vae = AutoencoderKL(VaeConfig(**vcfg.to_dict()), weights.StateDictSource(ovae.state_dict(), DEV), DEV, precision)
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler(), vae=vae)
    img = pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
               ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(),
               generator=torch.Generator().manual_seed(42), output_type="pt", **kw).images
    torch.cuda.synchronize()
    assert img.shape == ref_img.shape == (1, 3, 128, 128)
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    p = ov.psnr(img.cpu(), ref_img)
    assert p >= min_psnr, f"{precision}: PSNR {p:.1f} dB"
    arr = postprocess(torch.zeros(1, 3, 4, 4), "np")
    assert arr.shape == (1, 4, 4, 3) and float(arr.mean()) == 0.5


def test_pipeline_without_vae_rejects_image_output():
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0)
    usd, _ = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    pipe = InstantIRPipeline(UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV), DEV, "fp32"),
                             Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "fp32"), DDPMScheduler())
    inp = make_inputs(oc, B=1, h=32, w=32)
    with pytest.raises(ValueError):
        pipe(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
             pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
             ip_adapter_image_embeds=[inp["ip"]], num_inference_steps=1, preview_start=1.0, output_type="pt")


def test_full_size_sdxl_vae_decode_bf16_vs_fp32_check_mode():
    """SDXL VAE widths (128, 256, 512, 512), random-init: the tcgen05 path agrees with the fp32 check mode (itself
    oracle-exact at small size) at a 64² latent (512² image: 4096-token mid attention), image PSNR >= 40 dB; at the
    full 128² latent (1024² image, 16384-token attention in two query chunks) the decode is finite and run-to-run
    bit-identical."""
    cfg = VaeConfig()
    shapes = vae_decoder_param_shapes(cfg)
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(3)).to(DEV)
    imgs = {}
    for prec in ("fp32", "bf16"):
        vae = AutoencoderKL(cfg, weights.RandomSource(shapes, DEV, seed=7), DEV, prec)
        imgs[prec] = vae.decode(z).sample
        assert imgs[prec].shape == (1, 3, 512, 512) and torch.isfinite(imgs[prec]).all()
    assert rel_l2(imgs["bf16"], imgs["fp32"]) < 2e-2
    assert ov.psnr(postprocess(imgs["bf16"]).cpu(), postprocess(imgs["fp32"]).cpu()) >= 40.0
    z2 = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(4)).to(DEV)
    a = vae.decode(z2).sample.clone()
    b = vae.decode(z2).sample
    torch.cuda.synchronize()
    assert a.shape == (1, 3, 1024, 1024) and torch.isfinite(a).all() and torch.equal(a, b)


# ------------------------------------------------------------------------------------------ encode
def _full_oracle_vae(cfg, seed=81):
    vae = ov.AutoencoderKL(cfg)
    seeded_init(vae, seed)
    return vae.eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_encoder_vs_reference_run_vector(precision, tol):
    """product Encoder vs the reference's own Encoder.forward output (tests/golden/make_golden_vae.py): stride-2
    convs padded bottom/right only, one-head mid attention on a 8x6 grid, 8-channel conv_out"""
    from instantir_b200.vae import vae_param_shapes

    g = torch.load(os.path.join(G, "vae_encoder.pt"))
    cfg = ov.VaeConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["cfg"].items()})
    ovae = _full_oracle_vae(cfg)
    seeded_init(ovae.encoder, g["seed"])
    L2 = 2 * cfg.latent_channels
    with torch.no_grad():
        ovae.quant_conv.weight.copy_(torch.eye(L2).reshape(L2, L2, 1, 1))
        ovae.quant_conv.bias.zero_()
    sd = ovae.state_dict()
    shapes = vae_param_shapes(VaeConfig(**cfg.to_dict()))
    assert set(shapes) == set(sd) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    vae = AutoencoderKL(VaeConfig(**cfg.to_dict()), weights.StateDictSource(sd, DEV), DEV, precision)
    dist = vae.encode(g["x"].to(DEV)).latent_dist
    torch.cuda.synchronize()
    assert dist.parameters.shape == g["out"].shape
    assert rel_l2(dist.parameters, g["out"]) < tol


def test_gaussian_distribution_vs_reference_run_vector():
    from instantir_b200 import ops
    from instantir_b200.vae import DiagonalGaussianDistribution

    g = torch.load(os.path.join(G, "vae_encoder.pt"))
    m, noise = g["moments"].to(DEV), g["noise"].to(DEV).contiguous()
    m[0, 4:, 0, 0] = 50.0   # logvar above the clamp at 20
    m[1, 4:, 1, 1] = -80.0  # and below the clamp at -30
    mean, logvar = m.cpu().chunk(2, dim=1)
    ref = mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * g["noise"]
    out = ops.gaussian_sample(m.contiguous(), noise, torch.empty_like(noise), scale=1.0)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-6
    assert rel_l2(ops.gaussian_sample(g["moments"].to(DEV), noise, torch.empty_like(noise)), g["sample"]) < 1e-6
    d = DiagonalGaussianDistribution(g["moments"].to(DEV))
    assert torch.equal(d.mode().cpu(), g["mode"])
    s = d.sample(torch.Generator().manual_seed(9), scale=0.5)
    n9 = torch.randn(g["noise"].shape, generator=torch.Generator().manual_seed(9))
    assert rel_l2(s, 0.5 * ov.gaussian_sample(g["moments"], n9)) < 1e-6


@pytest.mark.parametrize("precision,min_psnr", [("fp32", 60.0), ("fp16", 40.0), pytest.param("bf16", 40.0, marks=_BF16_XFAIL)])
def test_pipeline_from_pixels_to_pixels_psnr_vs_oracle(precision, min_psnr):
    """the whole f1 row around the loop: a 3-channel image in [-1, 1] -> vae.encode -> sample * scaling_factor ->
    2 denoising steps (previewer off) -> vae.decode -> image, against the oracle doing the same with the same draws"""
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0)
    vcfg = ov.tiny_vae()
    ovae = _full_oracle_vae(vcfg)
    inp = make_inputs(oc, B=1, h=32, w=32)
    gi = torch.Generator().manual_seed(77)
    img_in = (torch.rand(1, 3, 128, 128, generator=gi) * 2 - 1) * 0.8
    kw = dict(num_inference_steps=2, guidance_scale=7.0, preview_start=1.0)
    g = torch.Generator().manual_seed(42)
    lq = ov.image_to_latents(ovae, img_in, torch.randn(1, 4, 32, 32, generator=g))
    ref_lat = opipe.restore_latents(
        ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=lq,
        prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
        pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
        ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], generator=g, **kw)
    ref_img = ov.latents_to_image(ovae, ref_lat)
    usd, _ = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV), DEV, precision)
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, precision)
    vae = AutoencoderKL(VaeConfig(**vcfg.to_dict()), weights.StateDictSource(ovae.state_dict(), DEV), DEV, precision)
    lq_p = vae.encode(img_in.to(DEV)).latent_dist.sample(torch.Generator().manual_seed(42), scale=vcfg.scaling_factor)
    assert rel_l2(lq_p, lq) < (1e-4 if precision == "fp32" else 2e-2)
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler(), vae=vae)
    img = pipe(image=img_in, prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
               pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
               ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(),
               generator=torch.Generator().manual_seed(42), output_type="pt", **kw).images
    torch.cuda.synchronize()
    p = ov.psnr(img.cpu(), ref_img)
    assert p >= min_psnr, f"{precision}: PSNR {p:.1f} dB"
    with pytest.raises(TypeError):
        InstantIRPipeline(unet, agg, DDPMScheduler())(image=img_in, prompt_embeds=inp["prompt_embeds"],
                                                      pooled_prompt_embeds=inp["pooled_prompt_embeds"],
                                                      negative_prompt_embeds=inp["negative_prompt_embeds"],
                                                      negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                                                      ip_adapter_image_embeds=[inp["ip"]])


def test_preview_rows_are_decoded_and_step_callbacks_run():
    """save_preview_row with an image output decodes every stored preview latent (pipelines/sdxl_instantir.py:1706-1725);
    callback_on_step_end sees every step and may replace the latents (:1650-1658)."""
    oc = ocfg.tiny()
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
    vcfg = ov.tiny_vae()
    ovae = _oracle_vae(vcfg)
    inp = make_inputs(oc, B=1, h=32, w=32)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=8.0 / oc.lora_rank), DEV, "fp32")
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "fp32")
    vae = AutoencoderKL(VaeConfig(**vcfg.to_dict()), weights.StateDictSource(ovae.state_dict(), DEV), DEV, "fp32")
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler(), vae=vae)
    seen = []

    def cb(p, i, t, kw):
        seen.append((i, int(t), tuple(kw["latents"].shape)))
        return {"latents": kw["latents"] * 1.0}

    call = dict(image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(), num_inference_steps=2,
                guidance_scale=7.0, preview_start=0.0, save_preview_row=True)
    img, rows = pipe(generator=torch.Generator().manual_seed(42), output_type="pt", return_dict=False, callback_on_step_end=cb, **call)
    lat, lrows = pipe(generator=torch.Generator().manual_seed(42), output_type="latent", return_dict=False, **call)
    torch.cuda.synchronize()
    assert seen == [(0, 501, (1, 4, 32, 32)), (1, 1, (1, 4, 32, 32))]
    assert len(rows) == len(lrows) == 2 and rows[0].shape == (1, 3, 128, 128)
    for r, lz in zip(rows, lrows):
        assert torch.equal(r, postprocess(vae.decode(lz / vcfg.scaling_factor).sample))
    assert torch.equal(img, postprocess(vae.decode(lat / vcfg.scaling_factor).sample))
