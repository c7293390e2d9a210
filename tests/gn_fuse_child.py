"""Child process of tests/test_zz_gn_fuse_gpu.py and of bench.py's `experimental_gn_fuse` leg: exercises the OPT-IN fused
GroupNorm path (IIR_GN_FUSE=1: statistics accumulated by the producing GEMM / conv epilogue, iir_groupnorm_apply_sums) in a
process of its own, so that a fault in that not-yet-GPU-verified path cannot poison the CUDA context of the caller.

    python tests/gn_fuse_child.py kernels        per-kernel numerics of gn_sums + apply_sums against torch
    python tests/gn_fuse_child.py model          BASELINE config 1 full step, fused path vs the CPU oracle and vs the default path
    python tests/gn_fuse_child.py sdxl           one UNet + Aggregator step at SDXL widths (latent 32²): fused vs default path

Prints one JSON object on the last line of stdout: {"ok": bool, "checks": {name: value}, "error": str | null}."""
import json
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

DEV = "cuda"
S1, S2 = 2.0 ** 24, 2.0 ** 26  # csrc/common.cuh GN_S1_SCALE / GN_S2_SCALE


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def _sums_ref(out, n_img, groups):
    """(sum, sum of squares) per (sample, group) of a [M, N] fp32 tensor, in fp64"""
    M, N = out.shape
    t = out.double().view(n_img, M // n_img, groups, N // groups)
    return t.sum((1, 3)), (t * t).sum((1, 3))


def _check_sums(name, gn, out, n_img, groups, checks, tol_s=2e-3, tol_q=2e-4):
    s_ref, q_ref = _sums_ref(out, n_img, groups)
    s = gn[..., 0].double() / S1
    q = gn[..., 1].double() / S2
    # the sum is compared on the scale of the group's L2 norm, the sum of squares relatively
    e_s = float(((s - s_ref).abs() / (q_ref.sqrt() + 1.0)).max())
    e_q = float(((q - q_ref).abs() / q_ref.clamp_min(1e-12)).max())
    checks[name + ":sum_err"] = e_s
    checks[name + ":sumsq_rel_err"] = e_q
    assert e_s < tol_s and e_q < tol_q, (name, e_s, e_q)


def run_kernels(checks):
    from instantir_b200 import ops

    H16 = torch.float16
    # ---- linear GEMM, fp32 out + fp32 residual (transformer proj_out / resnet shortcut shape class), several tilings
    for (M, N, K, groups, n_img, bn, cluster) in [(256, 320, 128, 32, 2, 160, 1), (256, 320, 128, 32, 2, 256, 1), (512, 640, 192, 32, 2, 224, 2),
                                                   (2048, 1280, 1280, 32, 2, None, None), (384, 64, 64, 32, 3, 64, 1), (300, 128, 64, 32, 1, 96, 1)]:
        rows_ok = M % n_img == 0 and (M // n_img) % 32 == 0
        a = rnd(M, K, seed=1, dtype=H16)
        w = rnd(N, K, seed=2, scale=K ** -0.5, dtype=H16)
        bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
        out = torch.full((M, N), float("nan"), device=DEV)
        gn = torch.zeros(n_img, groups, 2, device=DEV, dtype=torch.int64)
        name = f"lin:{M}x{N}x{K}:bn{bn}:cl{cluster}"
        if not rows_ok:
            try:
                ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=res, bn=bn, cluster=cluster, gn=gn, rows_per_sample=M // n_img)
                raise AssertionError("an ineligible launch was accepted: " + name)
            except ops._lib.IIRError:
                checks[name + ":rejected"] = True
            continue
        ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=res, bn=bn, cluster=cluster, gn=gn, rows_per_sample=M // n_img)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias + res
        checks[name + ":out"] = rel_l2(out, ref)
        assert rel_l2(out, ref) < 3e-3, name
        _check_sums(name, gn, out, n_img, groups, checks)
        # determinism: integer accumulation must not depend on arrival order
        gn2 = torch.zeros_like(gn)
        ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=res, bn=bn, cluster=cluster, gn=gn2, rows_per_sample=M // n_img)
        torch.cuda.synchronize()
        assert torch.equal(gn, gn2), name + ": sums differ between two runs"
    # ---- implicit-GEMM conv, 16-bit out + time-embedding row vector (conv1) and fp32 out + residual (conv2)
    for (n, Hh, Ww, Cin, Cout, groups, bn) in [(2, 16, 16, 64, 320, 32, 160), (2, 32, 32, 64, 640, 32, None), (1, 8, 16, 128, 64, 32, 64),
                                               (2, 64, 64, 64, 320, 32, None)]:
        x = rnd(n, Hh, Ww, Cin, seed=1, dtype=H16)
        w = rnd(Cout, 9 * Cin, seed=2, scale=(9 * Cin) ** -0.5, dtype=H16)
        bias, rv = rnd(Cout, seed=3), rnd(n, Cout, seed=4)
        M = n * Hh * Ww
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2), bias, padding=1)
        ref = ref.permute(0, 2, 3, 1).reshape(M, Cout) + rv.repeat_interleave(Hh * Ww, 0)
        for odt, res in ((H16, None), (torch.float32, rnd(M, Cout, seed=5))):
            out = torch.empty(M, Cout, device=DEV, dtype=odt)
            gn = torch.zeros(n, groups, 2, device=DEV, dtype=torch.int64)
            name = f"conv:{n}x{Hh}x{Ww}x{Cin}->{Cout}:{'h16' if odt == H16 else 'f32+res'}"
            ops.gemm(x, w, out, M=M, N=Cout, K=9 * Cin, bias=bias, rowvec=rv, rows_per_sample=Hh * Ww, residual=res, bn=bn,
                     conv=dict(n_img=n, H=Hh, W=Ww, Cin=Cin), gn=gn)
            torch.cuda.synchronize()
            want = ref + (res if res is not None else 0)
            checks[name + ":out"] = rel_l2(out, want)
            assert rel_l2(out, want) < 4e-3, name
            if odt == torch.float32:
                _check_sums(name, gn, out, n, groups, checks)  # against the kernel's own fp32 output: tight
            else:
                # the statistics are those of the fp32 values BEFORE the 16-bit rounding of the store: compared with the
                # fp32 torch convolution of the same 16-bit operands (tensor-core vs torch accumulation order: loose)
                _check_sums(name, gn, want, n, groups, checks, tol_s=2e-2, tol_q=2e-2)
            # ---- consumer: one-pass GroupNorm from the sums vs the two-kernel GroupNorm on the same tensor
            g_, b_ = rnd(Cout, seed=6), rnd(Cout, seed=7)
            for silu in (False, True):
                y1 = torch.empty(M, Cout, device=DEV, dtype=H16)
                y2 = torch.empty(M, Cout, device=DEV, dtype=H16)
                ops.groupnorm_apply_sums(out, g_, b_, gn, y1, n_img=n, HW=Hh * Ww, C=Cout, groups=groups, eps=1e-5, silu=silu)
                ops.groupnorm(out, g_, b_, y2, n_img=n, HW=Hh * Ww, C=Cout, groups=groups, eps=1e-5, silu=silu)
                torch.cuda.synchronize()
                yr = F.group_norm(out.float().view(n, Hh * Ww, Cout).permute(0, 2, 1), groups, g_, b_, 1e-5).permute(0, 2, 1).reshape(M, Cout)
                yr = F.silu(yr) if silu else yr
                checks[name + f":apply(silu={int(silu)})_vs_torch"] = rel_l2(y1, yr)
                checks[name + f":apply(silu={int(silu)})_vs_two_kernel"] = rel_l2(y1, y2)
                assert rel_l2(y1, yr) < 3e-3 and rel_l2(y1, y2) < 2e-3, name
    # ---- memset
    t = torch.ones(1000, device=DEV, dtype=torch.int64)
    ops.memset_zero(t)
    torch.cuda.synchronize()
    assert not bool(t.any())


def _models(oc, precision, fuse, usd, ulora, asd, alpha):
    from instantir_b200 import config as pcfg, weights
    from instantir_b200.aggregator import Aggregator
    from instantir_b200.unet import UNet2DConditionModel

    os.environ["IIR_GN_FUSE"] = "1" if fuse else "0"
    pc = pcfg.ModelConfig(**oc.to_dict())
    unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=(alpha / oc.lora_rank) if ulora else 1.0), DEV, precision)
    agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, precision)
    assert unet.rt.gn_fuse == bool(fuse) and agg.rt.gn_fuse == bool(fuse)
    return unet, agg


def run_model(checks):
    """BASELINE config 1 (scaled-down UNet + aggregator + IP-adapter + LoRA previewer), 2 steps, CFG 7, fp16"""
    from _util import build_oracle, export_state, make_inputs
    from instantir_b200 import ops
    from instantir_b200.pipeline import InstantIRPipeline
    from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
    from oracle import config as ocfg, pipeline as opipe, schedulers as osched

    torch.set_grad_enabled(False)
    oc = ocfg.tiny()
    alpha = 8.0
    ounet, oagg = build_oracle(oc, seed=0, lora_alpha=alpha)
    inp = make_inputs(oc, B=1, h=32, w=32)
    common = dict(num_inference_steps=2, guidance_scale=7.0, preview_start=0.0)
    rec_o = {}
    opipe.restore_latents(ounet, oagg, osched.DDPMScheduler(), osched.LCMSingleStepScheduler(), image=inp["image"],
                          prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                          pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                          ip_image_embeds=inp["ip"], add_time_ids=inp["time_ids"], generator=torch.Generator().manual_seed(42), record=rec_o, **common)
    usd, ulora = export_state(ounet)
    asd, _ = export_state(oagg)
    outs = {}
    for fuse in (0, 1):
        unet, agg = _models(oc, "fp16", fuse, usd, ulora, asd, alpha)
        for graph in (False, True):
            rec = {}
            n0 = ops._lib.launch_count()
            InstantIRPipeline(unet, agg, DDPMScheduler())(
                image=inp["image"], prompt_embeds=inp["prompt_embeds"], negative_prompt_embeds=inp["negative_prompt_embeds"],
                pooled_prompt_embeds=inp["pooled_prompt_embeds"], negative_pooled_prompt_embeds=inp["negative_pooled_prompt_embeds"],
                ip_adapter_image_embeds=[inp["ip"]], previewer_scheduler=LCMSingleStepScheduler(),
                generator=torch.Generator().manual_seed(42), use_cuda_graph=graph, record=rec, **common)
            torch.cuda.synchronize()
            outs[(fuse, graph)] = rec["latents"]
            checks[f"config1:fuse{fuse}:graph{int(graph)}:launches"] = int(ops._lib.launch_count() - n0)
            for i, (a, b) in enumerate(zip(rec["latents"], rec_o["latents"])):
                e = rel_l2(a, b)
                checks[f"config1:fuse{fuse}:graph{int(graph)}:step{i}_vs_oracle"] = e
                assert e < 1e-2, (fuse, graph, i, e)
    for graph in (False, True):
        for i, (a, b) in enumerate(zip(outs[(1, graph)], outs[(0, graph)])):
            checks[f"config1:fused_vs_default:graph{int(graph)}:step{i}"] = rel_l2(a, b)
    assert checks["config1:fuse1:graph0:launches"] < checks["config1:fuse0:graph0:launches"], "the fused path did not remove any launch"


def run_sdxl(checks):
    """one UNet + Aggregator step at full SDXL widths, latent 32² (levels 32², 16², 8²), fused vs default path on identical
    random-init weights; previewer off"""
    from instantir_b200 import config as pcfg, ops, weights
    from instantir_b200.aggregator import Aggregator
    from instantir_b200.pipeline import InstantIRPipeline
    from instantir_b200.schedulers import DDPMScheduler, LCMSingleStepScheduler
    from instantir_b200.unet import UNet2DConditionModel

    torch.set_grad_enabled(False)
    cfg = pcfg.sdxl()
    g = torch.Generator().manual_seed(5)
    B, lat = 1, 32
    inp = dict(image=torch.randn(B, 4, lat, lat, generator=g) * 0.8, prompt_embeds=torch.randn(B, 77, 2048, generator=g),
               negative_prompt_embeds=torch.randn(B, 77, 2048, generator=g), pooled_prompt_embeds=torch.randn(B, 1280, generator=g),
               negative_pooled_prompt_embeds=torch.randn(B, 1280, generator=g),
               ip_adapter_image_embeds=[torch.stack([0.3 * torch.randn(B, 257, 1024, generator=g), torch.randn(B, 257, 1024, generator=g)])])
    res = {}
    for fuse in (0, 1):
        os.environ["IIR_GN_FUSE"] = str(fuse)
        unet = UNet2DConditionModel(cfg, weights.RandomSource(weights.unet_param_shapes(cfg, adapter=True), DEV, seed=0), DEV, "fp16")
        agg = Aggregator(cfg, weights.RandomSource(weights.aggregator_param_shapes(cfg), DEV, seed=1), DEV, "fp16")
        rec = {}
        n0 = ops._lib.launch_count()
        loop = InstantIRPipeline(unet, agg, DDPMScheduler())(
            **inp, num_inference_steps=30, guidance_scale=7.0, previewer_scheduler=LCMSingleStepScheduler(), preview_start=1.0,
            generator=torch.Generator().manual_seed(42), record=rec, prepare_only=True, use_cuda_graph=False)
        loop.step(0)
        torch.cuda.synchronize()
        res[fuse] = (rec["latents"][0].clone(), rec["pred_x0"][0].clone())
        checks[f"sdxl32:fuse{fuse}:launches"] = int(ops._lib.launch_count() - n0)
        del unet, agg, loop
        torch.cuda.empty_cache()
    checks["sdxl32:latents_fused_vs_default"] = rel_l2(res[1][0], res[0][0])
    checks["sdxl32:pred_x0_fused_vs_default"] = rel_l2(res[1][1], res[0][1])
    assert torch.isfinite(res[1][0]).all()
    # both paths normalise the same tensors; they differ by where the statistics are rounded (fp32 values vs the stored
    # 16-bit values), i.e. by far less than the fp16-vs-oracle error budget of 1e-2
    assert checks["sdxl32:latents_fused_vs_default"] < 2e-3
    assert checks["sdxl32:fuse1:launches"] < checks["sdxl32:fuse0:launches"]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "kernels"
    checks, err = {}, None
    try:
        assert torch.cuda.is_available(), "needs a CUDA device"
        {"kernels": run_kernels, "model": run_model, "sdxl": run_sdxl}[what](checks)
    except BaseException as e:  # noqa: BLE001 — reported to the parent, which decides
        err = f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}"
    print(json.dumps({"ok": err is None, "what": what, "checks": checks, "error": err}))
    sys.exit(0 if err is None else 1)


if __name__ == "__main__":
    main()
