"""CPU tests of the host-side logic (no GPU, no compute calls into the library)."""
import ast
import os
import subprocess
import sys

import pytest
import torch

from _util import build_oracle, export_state  # noqa: F401
from instantir_b200 import config as pcfg
from instantir_b200 import parallel, pipeline, schedulers, weights
from oracle import config as ocfg
from oracle import model as om
from oracle import pipeline as opipe
from oracle import schedulers as osched

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_shapes(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


@pytest.mark.parametrize("name", ["tiny", "sdxl"])
def test_param_inventory_matches_oracle_modules(name):
    """the product's key/shape inventory == the oracle modules' state dicts (== reference key layout)."""
    oc, pc = getattr(ocfg, name)(), getattr(pcfg, name)()
    assert oc.to_dict() == pc.to_dict()
    with torch.device("meta"):
        unet = om.load_adapter(om.UNet2DConditionModel(oc))
        agg = om.Aggregator(oc)
        om.remove_attn2(agg)
    assert _oracle_shapes(unet) == dict(weights.unet_param_shapes(pc, adapter=True))
    assert _oracle_shapes(agg) == dict(weights.aggregator_param_shapes(pc))
    if name == "sdxl":
        n = sum(torch.Size(s).numel() for k, s in weights.unet_param_shapes(pc, adapter=False).items())
        assert n == 2_567_463_684
        assert sum(torch.Size(s).numel() for s in weights.aggregator_param_shapes(pc).values()) == 1_004_809_280


def test_lora_targets_match_oracle_peft_rule():
    from oracle import lora as olora

    oc = ocfg.tiny()
    unet = om.load_adapter(om.UNet2DConditionModel(oc))
    wrapped = sorted(olora.add_previewer_lora(unet, oc.lora_rank, 8.0))
    shapes = weights.unet_param_shapes(pcfg.tiny(), adapter=True)
    assert wrapped == weights.lora_targets(shapes)
    _, lora = export_state(unet)
    assert {k: tuple(v.shape) for k, v in lora.items()} == dict(weights.lora_param_shapes(pcfg.tiny(), shapes))
    # not hit: conv_in/out, embeddings, norms, attn2.to_k/to_v, Resampler FF
    assert not any(t.endswith(("attn2.to_k", "attn2.to_v", "conv_in", "conv_out")) or "embedding" in t for t in wrapped)


def test_merge_lora_equals_two_path_forward():
    torch.manual_seed(0)
    w, a, b = torch.randn(6, 5), torch.randn(3, 5), torch.randn(6, 3)
    x = torch.randn(4, 5)
    m = weights.merge_lora(w, (a, b, 0.5))
    assert torch.allclose(x @ m.t(), x @ w.t() + 0.5 * (x @ a.t()) @ b.t(), atol=1e-5)
    wc, ac, bc = torch.randn(6, 5, 3, 3), torch.randn(3, 5, 3, 3), torch.randn(6, 3, 1, 1)
    xc = torch.randn(2, 5, 7, 7)
    F = torch.nn.functional
    ref = F.conv2d(xc, wc, padding=1) + 0.5 * F.conv2d(F.conv2d(xc, ac, padding=1), bc)
    assert torch.allclose(F.conv2d(xc, weights.merge_lora(wc, (ac, bc, 0.5)), padding=1), ref, atol=1e-4)


def test_step_masks_and_timesteps_match_oracle():
    for args in [(30, 0.0, 1.0, 0.0, 1.0), (30, 1.0, 1.0, 0.0, 1.0), (30, 0.0, 1.0, 0.0, 0.6), (2, 0.5, 1.0, 0.0, 1.0)]:
        assert pipeline.step_masks(*args) == opipe.step_masks(*args)
    keep, prev = pipeline.step_masks(30, 0.0, 1.0, 0.0, 0.6)
    assert sum(keep) == 18 and sum(prev) == 30  # config 4: 18 full + 12 UNet-only steps
    s, o = schedulers.DDPMScheduler(), osched.DDPMScheduler()
    for n in (30, 2, 50):
        s.set_timesteps(n)
        o.set_timesteps(n)
        assert s.timesteps.tolist() == o.timesteps.tolist()
        for t in s.timesteps:
            for x, y in zip(s.coefficients(t), o.coefficients(t)):
                assert float(x) == float(y)
    s.set_timesteps(timesteps=[900, 500, 100])
    o.set_timesteps(timesteps=[900, 500, 100])
    assert s.previous_timestep(500) == 100 and s.previous_timestep(100) == -1
    assert [float(v) for v in s.coefficients(100)] == [float(v) for v in o.coefficients(torch.tensor(100))]
    with pytest.raises(ValueError):
        s.set_timesteps(timesteps=[100, 500])
    lcm, olcm = schedulers.LCMSingleStepScheduler(), osched.LCMSingleStepScheduler()
    assert torch.equal(lcm.alphas_cumprod, olcm.alphas_cumprod)
    for t in (958, 501, 1, 0):
        a = lcm.get_scalings_for_boundary_condition_discrete(torch.tensor(t))
        b = olcm.get_scalings_for_boundary_condition_discrete(torch.tensor(t))
        assert float(a[0]) == float(b[0]) and float(a[1]) == float(b[1])


def test_partition_covers_every_image_once():
    for n, ws, cfgp in [(64, 8, True), (64, 8, False), (5, 4, True), (3, 8, False), (1, 2, True)]:
        seen = []
        for r in range(ws):
            sl, br = parallel.partition(n, ws, r, cfgp)
            assert br == (r % 2 if cfgp else None)
            if not cfgp or br == 0:
                seen += list(range(n))[sl]
            if cfgp:
                assert sl == parallel.partition(n, ws, r ^ 1, cfgp)[0]  # both ranks of a pair share images
        assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        parallel.partition(4, 3, 0, True)


def test_random_source_is_deterministic_and_nonzero():
    shapes = weights.unet_param_shapes(pcfg.tiny())
    lshapes = weights.lora_param_shapes(pcfg.tiny(), shapes)
    a = weights.RandomSource(shapes, "cpu", seed=3, lora_shapes=lshapes)
    b = weights.RandomSource(shapes, "cpu", seed=3, lora_shapes=lshapes)
    k = "down_blocks.1.attentions.0.transformer_blocks.0.attn2.processor.ln_k_ip.linear.weight"
    assert torch.equal(a.get(k), b.get(k)) and float(a.get(k).abs().sum()) > 0
    la = a.get_lora("down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q")
    assert la is not None and float(la[1].abs().sum()) > 0
    assert a.get_lora("down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k") is None


_GLOO = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from instantir_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
cp = parallel.CFGParallel()
assert cp.branch == dist.get_rank()
eps = torch.full((3, 4, 2, 2), float(cp.branch + 1))
out = cp.gather_branches(eps)
assert out.shape == (6, 4, 2, 2) and float(out[:3].mean()) == 1.0 and float(out[3:].mean()) == 2.0
# both ranks then run the same CFG combine with identical noise -> identical latents
g = torch.Generator().manual_seed(42)
z = torch.randn(3, 4, 2, 2, generator=g)
lat = out[:3] + 7.0 * (out[3:] - out[:3]) + z
ref = [torch.empty_like(lat) for _ in range(2)]
dist.all_gather(ref, lat)
assert torch.equal(ref[0], ref[1])
# ranks that arrive with DIFFERENT generator states still step identical noise: the pair leader's draw wins
mine_z = torch.randn(3, 4, 2, 2, generator=torch.Generator().manual_seed(100 + dist.get_rank()))
want_z = torch.randn(3, 4, 2, 2, generator=torch.Generator().manual_seed(100))
assert torch.equal(cp.broadcast_from_leader(mine_z), want_z)
sl, br = parallel.partition(5, 2, dist.get_rank(), cfg_parallel=False)
full = torch.randn(5, 4, generator=torch.Generator().manual_seed(1))
mine = parallel.draw_shared_noise((5, 4), torch.Generator().manual_seed(1), "cpu", sl)
assert torch.equal(mine, full[sl])
dist.destroy_process_group()
print("ok")
"""


def test_cfg_parallel_gather_world_size_2_gloo(tmp_path):
    """the N>1 path on CPU: 2 ranks, gloo, one all-gather of eps, identical latents on both ranks."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "gloo_pair.py"
    script.write_text(_GLOO.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_product_never_imports_oracle_or_reference():
    """the product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "instantir_b200")
    for fn in os.listdir(pkg):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                mods = [node.module]
            for m in mods:
                assert not m.split(".")[0] in ("oracle", "diffusers", "peft", "triton"), (fn, m)
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        for line in open(os.path.join(pkg, "csrc", fn)):
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line and "reference" not in line, (fn, line)


def test_pipeline_rejects_out_of_scope_arguments():
    from types import SimpleNamespace

    p = pipeline.InstantIRPipeline.__new__(pipeline.InstantIRPipeline)
    p.text_encoder = p.text_encoder_2 = p.tokenizer = p.tokenizer_2 = p.image_encoder = p.feature_extractor = None
    p.config = SimpleNamespace(force_zeros_for_empty_prompt=True)
    with pytest.raises(ValueError, match="text_encoder_2"):   # a string prompt needs the encoders + tokenizers
        pipeline.InstantIRPipeline.__call__(p, prompt="a photo")
    with pytest.raises(ValueError, match="Cannot forward both"):
        pipeline.InstantIRPipeline.__call__(p, prompt="a photo", prompt_embeds=torch.zeros(1, 77, 8))
    with pytest.raises(NotImplementedError):
        pipeline.InstantIRPipeline.__call__(p, multistep_restore=True)
    with pytest.raises(NotImplementedError):
        pipeline.InstantIRPipeline.__call__(p, reference_latents=torch.zeros(1, 4, 8, 8), agg_ahead=True)


def test_from_unet_weight_source_matches_reference_from_unet():
    """Aggregator.from_unet (module/aggregator.py:503-578, used at pipelines/sdxl_instantir.py:320-322): every trunk tensor
    (conv_in, ref_conv_in <- the UNet's conv_in, time / add embeddings, down + mid blocks minus attn2 / norm2) is the UNet's
    own, the SFT heads get a bounded default Conv2d initialisation and every head closes with a ZERO 1x1 convolution — the
    product's weight source against the oracle's from_unet + remove_attn2 on the same seeded UNet."""
    from instantir_b200.aggregator import _FromUNetSource, _OverlaySource

    oc = ocfg.tiny()
    ounet, _ = build_oracle(oc, seed=0)
    usd, _ = export_state(ounet)
    pc = pcfg.ModelConfig(**oc.to_dict())
    fs = _FromUNetSource(pc, weights.StateDictSource(usd, "cpu"), "cpu", seed=0)
    oagg = om.Aggregator.from_unet(ounet)
    om.remove_attn2(oagg)
    osd = oagg.state_dict()
    shapes = weights.aggregator_param_shapes(pc)
    assert set(shapes) == set(osd)
    n_trunk = n_zero = n_init = 0
    for k, shp in shapes.items():
        t = fs.get(k)
        assert tuple(t.shape) == tuple(shp) and fs.has(k), k
        if not k.startswith("controlnet_"):
            assert torch.equal(t.float(), osd[k].float()), k
            n_trunk += 1
        elif k.startswith("controlnet_mid_block.1.") or k.split(".")[2] == "1":
            assert not t.any(), k  # zero_module (module/aggregator.py:980-983): residuals are exactly zero before load_state_dict
            n_zero += 1
        else:
            w = shapes[k.rsplit(".", 1)[0] + ".weight"]
            bound = (w[1] * w[2] * w[3]) ** -0.5
            assert float(t.abs().max()) <= bound and float(t.abs().max()) > 0.5 * bound, k  # torch's default Conv2d init range
            assert torch.equal(t, fs.get(k)), k                                              # and reproducible
            n_init += 1
    assert n_trunk > 100 and n_zero == 2 * 10 and n_init == 6 * 10  # 9 + 1 heads: (mlp_shared, mul, add) x (w, b) + zero conv (w, b)
    assert torch.equal(fs.get("ref_conv_in.weight"), fs.get("conv_in.weight"))
    with pytest.raises(KeyError):
        fs.get("up_blocks.0.resnets.0.conv1.weight")
    # load_state_dict(strict=False): keys of the new state dict win, absent keys keep their current tensors
    ov = _OverlaySource({"conv_in.bias": torch.full_like(osd["conv_in.bias"], 3.0)}, fs, "cpu")
    assert float(ov.get("conv_in.bias")[0]) == 3.0 and torch.equal(ov.get("conv_in.weight"), fs.get("conv_in.weight"))
    assert ov.has("conv_in.bias") and ov.has("mid_block.resnets.0.conv1.weight") and not ov.has("nope")


def test_aggregator_load_state_dict_validates_like_torch():
    """infer.py:142-144: `aggregator.load_state_dict(torch.load(aggregator.pt))`.  The key / shape validation runs before any
    device work, so it is checked here on a stand-in object: strict=True raises on missing, unexpected and mis-shaped keys."""
    from types import SimpleNamespace

    from instantir_b200.aggregator import Aggregator

    pc = pcfg.tiny()
    shapes = weights.aggregator_param_shapes(pc)
    stub = SimpleNamespace(state_dict_keys=lambda: shapes)
    good = {k: torch.zeros(s) for k, s in shapes.items()}
    missing = dict(good)
    missing.pop("conv_in.weight")
    with pytest.raises(RuntimeError, match="missing"):
        Aggregator.load_state_dict(stub, missing)
    with pytest.raises(RuntimeError, match="unexpected"):
        Aggregator.load_state_dict(stub, dict(good, extra=torch.zeros(1)))
    with pytest.raises(RuntimeError, match="size mismatch"):
        Aggregator.load_state_dict(stub, dict(good, **{"conv_in.bias": torch.zeros(3)}))
