"""CPU test of the host code of the widening rows f1 (VAE, instantir_b200/vae.py) and f2 (CLIP text / DINOv2 encoders,
instantir_b200/encoders.py): the real product objects with torch emulations of the kernel contracts (tests/_ops_emulation.py)
against the reference-run / transformers-run vectors and the oracle — the same scenarios as tests/test_vae_gpu.py and
tests/test_encoders_gpu.py, on the box without a GPU.  The kernels are checked by those GPU tests."""
import os

import pytest
import torch

import _ops_emulation
from _util import rel_l2
from seeding import seeded_init

from instantir_b200 import weights
from oracle import encoders as oe
from oracle import vae as ov

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"fp32": 1e-4, "fp16": 4e-3}
torch.set_grad_enabled(False)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_vae_decoder_and_encoder_host_code(monkeypatch, precision):
    from instantir_b200.vae import AutoencoderKL, VaeConfig

    _ops_emulation.install(monkeypatch.setattr)
    g = torch.load(os.path.join(G, "vae_decoder.pt"))
    cfg = ov.VaeConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["cfg"].items()})
    odec = seeded_init(ov.Decoder(cfg), g["seed"])
    sd = {"decoder." + k: v for k, v in odec.state_dict().items()}
    L = cfg.latent_channels
    sd["post_quant_conv.weight"] = torch.eye(L).reshape(L, L, 1, 1)
    sd["post_quant_conv.bias"] = torch.zeros(L)
    vae = AutoencoderKL(VaeConfig(**cfg.to_dict()), weights.StateDictSource(sd, "cpu"), "cpu", precision)
    out = vae.decode(g["z"]).sample
    assert out.shape == g["out"].shape and rel_l2(out, g["out"]) < TOL[precision]   # the reference's own Decoder.forward
    # encoder against the reference's own Encoder.forward (stride-2 convs padded bottom/right only, one-head mid attention)
    ge = torch.load(os.path.join(G, "vae_encoder.pt"))
    ecfg = ov.VaeConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in ge["cfg"].items()})
    ovae = seeded_init(ov.AutoencoderKL(ecfg), 72).eval()
    seeded_init(ovae.encoder, ge["seed"])
    L2 = 2 * ecfg.latent_channels
    ovae.quant_conv.weight.copy_(torch.eye(L2).reshape(L2, L2, 1, 1))
    ovae.quant_conv.bias.zero_()
    evae = AutoencoderKL(VaeConfig(**ecfg.to_dict()), weights.StateDictSource(ovae.state_dict(), "cpu"), "cpu", precision)
    dist = evae.encode(ge["x"]).latent_dist
    assert dist.parameters.shape == ge["out"].shape and rel_l2(dist.parameters, ge["out"]) < TOL[precision]
    s9 = dist.sample(torch.Generator().manual_seed(9), scale=0.5)
    n9 = torch.randn(s9.shape, generator=torch.Generator().manual_seed(9))
    assert rel_l2(s9, 0.5 * ov.gaussian_sample(dist.parameters.float(), n9)) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_text_and_image_encoder_host_code(monkeypatch, precision):
    from instantir_b200 import encoders as pe

    _ops_emulation.install(monkeypatch.setattr)
    g = torch.load(os.path.join(G, "encoders.pt"))
    tol = TOL[precision]
    for name in ("clip_l", "clip_g"):
        r = g[name]
        om = seeded_init(oe.CLIPTextModel(**r["cfg"]), r["seed"])
        cfg = pe.CLIPTextConfig(**{k: v for k, v in r["cfg"].items() if k != "projection_dim"}, projection_dim=r["cfg"].get("projection_dim") or 0)
        enc = pe.CLIPTextModel(cfg, weights.StateDictSource(om.state_dict(), "cpu"), "cpu", precision, with_projection=name == "clip_g")
        out = enc(g["ids"], output_hidden_states=True)
        assert rel_l2(out.hidden_states[-2], r["penultimate"]) < tol                 # transformers' own run
        assert rel_l2(out.last_hidden_state, r["last_hidden_state"]) < tol
        assert rel_l2(out.text_embeds if name == "clip_g" else out.pooler_output, r["text_embeds" if name == "clip_g" else "pooler_output"]) < tol
    r = g["dinov2"]
    om = seeded_init(oe.Dinov2Model(**r["cfg"]), r["seed"])
    enc = pe.Dinov2Model(pe.Dinov2Config(**r["cfg"]), weights.StateDictSource(om.state_dict(), "cpu"), "cpu", precision)
    assert rel_l2(enc(r["x70"]).last_hidden_state, r["out70"]) < tol
    assert rel_l2(enc(r["x42"]).last_hidden_state, r["out42"]) < tol                 # interpolated position embeddings
