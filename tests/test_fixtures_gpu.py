"""The CUDA product, called through its reference-shaped Python API (-> C ABI), DIRECTLY against the vectors
produced by running the reference's own files (tests/golden/make_golden.py, make_golden_misc.py):
attention processors (module/ip_adapter/attention_processor.py), LCM single-step scheduler
(schedulers/lcm_single_step_scheduler.py) and rescale_noise_cfg (pipelines/sdxl_instantir.py:181-192).
test_oracle_golden.py pins the CPU oracle on the same files; here no oracle arithmetic is involved at all —
oracle module classes are used only as seeded containers of the fixture's weights.

Tolerances: fp32 check mode 1e-4 (north star); fp16 3e-3, bf16 2e-2 for single-module outputs."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import rel_l2  # noqa: E402
from seeding import checksum, seeded_init  # noqa: E402

from instantir_b200 import ops, weights  # noqa: E402
from instantir_b200.attention_processor import AdaLayerNorm, Attention, AttnProcessor2_0, TA_IPAttnProcessor2_0  # noqa: E402
from instantir_b200.nn import Runtime  # noqa: E402
from instantir_b200.schedulers import LCMSingleStepScheduler  # noqa: E402
from oracle import model as om  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"
TOL = {"fp32": 1e-4, "fp16": 3e-3, "bf16": 2e-2}
torch.set_grad_enabled(False)


def _prefixed(module, prefix):
    return {prefix + k: v for k, v in module.state_dict().items()}


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_processors_vs_reference_run(precision):
    """TA_IPAttnProcessor2_0 (tuple and concatenated encoder_hidden_states), AttnProcessor2_0 and AdaLayerNorm of
    the product vs the outputs of the reference's classes (processors.pt)."""
    g = torch.load(os.path.join(G, "processors.pt"))
    d = g["dims"]
    oattn = seeded_init(om.Attention(d["C"], d["heads"], d["xdim"]), g["seeds"]["attn"])
    oproc = seeded_init(om.TA_IPAttnProcessor2_0(d["C"], d["xdim"], time_embedding_dim=d["tdim"], scale=d["scale"],
                                                 num_tokens=d["ntok"]), g["seeds"]["proc"])
    assert abs(checksum(oattn) - g["checksums"]["attn"]) <= 1e-9 * g["checksums"]["attn"]
    assert abs(checksum(oproc) - g["checksums"]["proc"]) <= 1e-9 * g["checksums"]["proc"]
    sd = {**_prefixed(oattn, "a."), **_prefixed(oproc, "a.processor.")}
    rt = Runtime(DEV, precision)
    src = weights.StateDictSource(sd, DEV)
    attn = Attention(rt, src, "a", d["C"], d["heads"], d["xdim"])
    proc = TA_IPAttnProcessor2_0(rt, src, "a.processor", d["C"], d["xdim"], time_embedding_dim=d["tdim"], scale=d["scale"],
                                 num_tokens=d["ntok"])
    hs, text, ip, temb = (g[k].to(DEV) for k in ("hs", "text", "ip", "temb"))
    out = proc(attn, hs, encoder_hidden_states=(text, [ip]), temb=temb)
    torch.cuda.synchronize()
    assert rel_l2(out, g["out_tuple"]) < TOL[precision]
    attn.ctx.invalidate()
    out = proc(attn, hs, encoder_hidden_states=torch.cat([text, ip], 1), temb=temb)
    torch.cuda.synchronize()
    assert rel_l2(out, g["out_concat"]) < TOL[precision]
    with pytest.raises(AssertionError):  # attention_processor.py:1102
        proc(attn, hs, encoder_hidden_states=(text, [ip]), temb=None)

    osa = seeded_init(om.Attention(d["C"], d["heads"]), g["seeds"]["self_attn"])
    sa = Attention(rt, weights.StateDictSource(_prefixed(osa, "s."), DEV), "s", d["C"], d["heads"])
    out = AttnProcessor2_0()(sa, hs, temb=temb)
    torch.cuda.synchronize()
    assert rel_l2(out, g["out_self"]) < TOL[precision]

    oada = seeded_init(om.AdaLayerNorm(d["C"], d["tdim"]), g["seeds"]["ada"])
    ada = AdaLayerNorm(rt, weights.StateDictSource(_prefixed(oada, "n."), DEV), "n", d["C"])
    out = ada(hs.reshape(-1, d["C"]).contiguous(), temb, rows_per_sample=hs.shape[1])
    torch.cuda.synchronize()
    assert rel_l2(out.view_as(hs), g["out_ada"]) < (1e-4 if precision == "fp32" else 4e-3 if precision == "fp16" else 1e-2)


def test_lcm_scheduler_vs_reference_run():
    """product LCMSingleStepScheduler (iir_lcm_step / iir_add_noise) vs the reference file's own outputs.  The kernel
    multiplies by 1/sqrt(abar) where the reference divides: last-bit differences, hence 1e-6 rather than bit equality."""
    g = torch.load(os.path.join(G, "lcm_scheduler.pt"))
    lcm = LCMSingleStepScheduler()
    assert torch.equal(lcm.alphas_cumprod, g["alphas_cumprod"])
    eps, x = g["eps"].to(DEV), g["x"].to(DEV)
    for t, want in g["steps"].items():
        got = lcm.step(eps, torch.tensor(t, dtype=torch.int64), x, return_dict=False)[0]
        torch.cuda.synchronize()
        assert got.dtype == torch.float32 and rel_l2(got, want) < 1e-6, t
    assert torch.equal(lcm.step(eps, 0, x, return_dict=False)[0].cpu(), g["x"])  # t = 0: c_skip = 1, c_out = 0
    # per-sample timesteps (lcm_single_step_scheduler.py:492-513 indexes alphas_cumprod per sample)
    noisy = lcm.add_noise(x, eps, torch.tensor([958, 1]))
    torch.cuda.synchronize()
    assert rel_l2(noisy, g["noisy"]) < 1e-6
    # the 16-bit eps the tensor-core UNet could hand over is accepted too
    got = lcm.step(eps.half(), 501, x, return_dict=False)[0]
    assert rel_l2(got, g["steps"][501]) < 2e-3


def test_rescale_noise_cfg_vs_reference_run():
    """iir_cfg_rescale vs the reference's rescale_noise_cfg executed verbatim (sums in fp64 here, fp32 there)"""
    g = torch.load(os.path.join(G, "rescale_noise_cfg.pt"))
    e_u, e_c = g["e_u"].to(DEV), g["e_c"].to(DEV)
    for phi, want in g["out"].items():
        out = ops.cfg_rescale(e_u, e_c, torch.empty_like(e_u), guidance=g["guidance"], rescale=phi)
        torch.cuda.synchronize()
        assert rel_l2(out, want) < 2e-6, phi
