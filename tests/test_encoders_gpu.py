"""SURVEY §8 row f2 on the GPU: the product's CLIP text encoders and DINOv2 image encoder (instantir_b200/encoders.py,
through the C ABI) against vectors produced by running `transformers` itself (tests/golden/encoders.pt) and against the
CPU oracle (oracle/encoders.py) on other inputs.  fp32 check mode <= 1e-4, fp16 <= 3e-3, bf16 <= 2e-2."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import rel_l2  # noqa: E402
from seeding import seeded_init  # noqa: E402

from instantir_b200 import encoders as pe  # noqa: E402
from instantir_b200 import weights  # noqa: E402
from oracle import encoders as oe  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
TOL = {"fp32": 1e-4, "fp16": 3e-3, "bf16": 2e-2}
torch.set_grad_enabled(False)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["clip_l", "clip_g"])
def test_clip_text_encoders_vs_transformers_run(name, precision):
    g = torch.load(os.path.join(G, "encoders.pt"))
    r = g[name]
    om = seeded_init(oe.CLIPTextModel(**r["cfg"]), r["seed"])
    cfg = pe.CLIPTextConfig(**{k: v for k, v in r["cfg"].items() if k != "projection_dim"}, projection_dim=r["cfg"].get("projection_dim") or 0)
    shapes = pe.clip_text_param_shapes(cfg, with_projection=name == "clip_g")
    sd = om.state_dict()
    assert set(shapes) == set(sd) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    enc = pe.CLIPTextModel(cfg, weights.StateDictSource(sd, DEV), DEV, precision, with_projection=name == "clip_g")
    out = enc(g["ids"], output_hidden_states=True)
    torch.cuda.synchronize()
    tol = TOL[precision]
    assert len(out.hidden_states) == r["n_hidden"]
    assert rel_l2(out.hidden_states[-2], r["penultimate"]) < tol
    assert rel_l2(out.last_hidden_state, r["last_hidden_state"]) < tol
    if name == "clip_l":
        assert rel_l2(out.pooler_output, r["pooler_output"]) < tol
        assert out[0] is out.last_hidden_state
    else:
        assert rel_l2(out.text_embeds, r["text_embeds"]) < tol
        assert out[0] is out.text_embeds
    # a batch of one, shorter than 77 tokens, first-EOS rule (eos_token_id != 2): against the oracle
    om.eos_token_id = 7
    ids = torch.randint(8, 990, (1, 40), generator=torch.Generator().manual_seed(3))
    ids[0, 11] = 7
    ids[0, 30] = 7
    ref = om(ids)
    enc.config.eos_token_id = 7
    got = enc(ids)
    torch.cuda.synchronize()
    assert rel_l2(got.pooler_output, ref["pooler_output"]) < tol
    assert rel_l2(got.hidden_states[-2], ref["hidden_states"][-2]) < tol


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_dinov2_vs_transformers_run(precision):
    g = torch.load(os.path.join(G, "encoders.pt"))
    r = g["dinov2"]
    om = seeded_init(oe.Dinov2Model(**r["cfg"]), r["seed"])
    cfg = pe.Dinov2Config(**r["cfg"])
    shapes = pe.dinov2_param_shapes(cfg)
    sd = om.state_dict()
    assert set(shapes) == set(sd) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    enc = pe.Dinov2Model(cfg, weights.StateDictSource(sd, DEV), DEV, precision)
    tol = TOL[precision]
    out70 = enc(r["x70"]).last_hidden_state
    out42 = enc(r["x42"]).last_hidden_state                    # 3 x 4 patch grid: interpolated position embeddings
    torch.cuda.synchronize()
    assert out70.shape == r["out70"].shape and rel_l2(out70, r["out70"]) < tol
    assert out42.shape == r["out42"].shape and rel_l2(out42, r["out42"]) < tol
    # the reference's pinned transformers 4.36.2 resamples with scale factors (+0.1): against the oracle's restatement
    ref = om(r["x42"], pos_mode="scale_0.1")
    got = enc(r["x42"], pos_mode="scale_0.1").last_hidden_state
    torch.cuda.synchronize()
    assert rel_l2(got, ref) < tol
    # the zero image of encode_image's uncond branch (pipelines/sdxl_instantir.py:660-664)
    z = torch.zeros_like(r["x70"][:1])
    assert rel_l2(enc(z).last_hidden_state, om(z)) < tol


def test_full_size_encoders_run_and_are_deterministic():
    """SDXL's real geometries (CLIP ViT-L/14 text, OpenCLIP bigG/14 text, DINOv2-L at 224²): shapes, finiteness, determinism"""
    ids = torch.randint(0, 49407, (2, 77), generator=torch.Generator().manual_seed(1))
    ids[:, -1] = 49407
    for cfg, proj in ((pe.clip_l(), False), (pe.clip_bigg(), True)):
        enc = pe.CLIPTextModel(cfg, weights.RandomSource(pe.clip_text_param_shapes(cfg, proj), DEV, seed=3), DEV, "fp16", with_projection=proj)
        a, b = enc(ids), enc(ids)
        torch.cuda.synchronize()
        assert a.hidden_states[-2].shape == (2, 77, cfg.hidden_size) and torch.isfinite(a.hidden_states[-2]).all()
        assert torch.equal(a.hidden_states[-2], b.hidden_states[-2]) and torch.equal(a[0], b[0])
        assert a[0].shape == ((2, cfg.projection_dim) if proj else (2, 77, cfg.hidden_size))
        del enc
    cfg = pe.Dinov2Config()
    enc = pe.Dinov2Model(cfg, weights.RandomSource(pe.dinov2_param_shapes(cfg), DEV, seed=4), DEV, "fp16")
    img = torch.rand(2, 3, 300, 400, generator=torch.Generator().manual_seed(2))
    x = pe.dinov2_preprocess(img)
    assert x.shape == (2, 3, 224, 224)
    out = enc(x).last_hidden_state
    torch.cuda.synchronize()
    assert out.shape == (2, 257, 1024) and torch.isfinite(out).all()
    assert torch.equal(out, enc(x).last_hidden_state)


def test_pipeline_encodes_prompts_and_ip_image_fp32():
    """encode_prompt / prepare_ip_adapter_image_embeds (pipelines/sdxl_instantir.py:400-729) against the oracle encoders, and the
    pipeline called with token ids + an IP image must equal the pipeline called with the embeddings those produce."""
    from _util import build_oracle, export_state, make_inputs

    from instantir_b200 import config as pcfg
    from instantir_b200.aggregator import Aggregator
    from instantir_b200.pipeline import InstantIRPipeline
    from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
    from instantir_b200.unet import UNet2DConditionModel
    from oracle import config as ocfg

    oc = ocfg.tiny()   # cross_attention_dim 256 = 128 + 128 (two text encoders), pooled 64, DINO tokens [33, 64]
    kw_l = dict(vocab_size=500, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2)
    o_l = seeded_init(oe.CLIPTextModel(**kw_l, hidden_act="quick_gelu"), 51)
    o_g = seeded_init(oe.CLIPTextModel(**kw_l, hidden_act="gelu", projection_dim=oc.pooled_dim), 52)
    d_kw = dict(hidden_size=64, num_hidden_layers=2, num_attention_heads=1, image_size=56, patch_size=14)
    o_d = seeded_init(oe.Dinov2Model(**d_kw), 53)
    enc_l = pe.CLIPTextModel(pe.CLIPTextConfig(**kw_l, hidden_act="quick_gelu", projection_dim=0), weights.StateDictSource(o_l.state_dict(), DEV), DEV, "fp32")
    enc_g = pe.CLIPTextModel(pe.CLIPTextConfig(**kw_l, hidden_act="gelu", projection_dim=oc.pooled_dim),
                             weights.StateDictSource(o_g.state_dict(), DEV), DEV, "fp32", with_projection=True)
    dino = pe.Dinov2Model(pe.Dinov2Config(**d_kw), weights.StateDictSource(o_d.state_dict(), DEV), DEV, "fp32")
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=8.0)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=8.0 / oc.lora_rank), DEV, "fp32")
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, "fp32")
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler(), text_encoder=enc_l, text_encoder_2=enc_g, image_encoder=dino)
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(3, 490, (1, 77), generator=g)
    ids[0, 30] = 499
    nids = torch.randint(3, 490, (1, 77), generator=g)
    nids[0, 12] = 499
    px = torch.randn(1, 3, 56, 56, generator=g)   # 4 x 4 patches + cls = 17 tokens ... the tiny UNet expects 33
    px = torch.randn(1, 3, 56, 112, generator=g)  # 4 x 8 patches + cls = 33 tokens of width 64
    pe_, npe, pp, npp = pipe.encode_prompt(ids, negative_prompt=nids)
    ref = [torch.cat([o_l(i)["hidden_states"][-2], o_g(i)["hidden_states"][-2]], -1) for i in (ids, nids)]
    assert rel_l2(pe_, ref[0]) < 1e-4 and rel_l2(npe, ref[1]) < 1e-4
    assert rel_l2(pp, o_g(ids)["text_embeds"]) < 1e-4 and rel_l2(npp, o_g(nids)["text_embeds"]) < 1e-4
    z_pe, z_npe, z_pp, z_npp = pipe.encode_prompt(ids)    # no negative prompt: zeros (force_zeros_for_empty_prompt, :552-555)
    assert float(z_npe.abs().sum()) == 0.0 and float(z_npp.abs().sum()) == 0.0 and torch.equal(z_pe, pe_)
    ipe = pipe.prepare_ip_adapter_image_embeds(px, None)[0]
    assert ipe.shape == (2, 1, 33, 64)
    assert rel_l2(ipe[1], o_d(px)) < 1e-4 and rel_l2(ipe[0], o_d(torch.zeros_like(px))) < 1e-4
    inp = make_inputs(oc, B=1, h=32, w=32)
    kw = dict(image=inp["image"], num_inference_steps=2, guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=0.0)
    a = pipe(prompt=ids, negative_prompt=nids, ip_adapter_image=px, generator=torch.Generator().manual_seed(42), **kw).images
    b = pipe(prompt_embeds=pe_, negative_prompt_embeds=npe, pooled_prompt_embeds=pp, negative_pooled_prompt_embeds=npp,
             ip_adapter_image_embeds=[ipe], generator=torch.Generator().manual_seed(42), **kw).images
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and torch.equal(a, b)
    with pytest.raises(ValueError, match="tokenizer"):
        pipe(prompt="a photo", ip_adapter_image=px, **kw)
