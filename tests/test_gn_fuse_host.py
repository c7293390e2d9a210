"""CPU test of the HOST side of the opt-in fused GroupNorm path: nn.Runtime's accumulator arena, FMap.gn, the
producers (ResnetBlock2D.conv1 / conv2, the stride-2 downsampler GEMM) and the consumer (GroupNorm.__call__), with the C-ABI
calls of instantir_b200.ops replaced by torch emulations that follow the kernels' contracts (fixed-point sums included).  It
checks the plumbing — which statistics reach which GroupNorm, fall-backs, arena reuse across forwards — not the kernels
(those run in tests/test_zz_gn_fuse_gpu.py on the driver's GPU)."""
import torch
import torch.nn.functional as F

from instantir_b200 import nn, ops

S1, S2 = 2.0 ** 24, 2.0 ** 26


class _Fake:
    """torch stand-ins for the ops the blocks under test launch"""

    def __init__(self):
        self.calls = []

    def gemm(self, a, w, out, *, M, N, K, lda=None, bias=None, rowvec=None, rows_per_sample=0, residual=None, act=0, conv=None,
             tc=True, gn=None, **kw):
        self.calls.append("gemm+gn" if gn is not None else "gemm")
        if conv is not None:
            x = a.float().view(conv["n_img"], conv["H"], conv["W"], conv["Cin"]).permute(0, 3, 1, 2)
            y = F.conv2d(x, w.float().view(N, 3, 3, conv["Cin"]).permute(0, 3, 1, 2), None, padding=1).permute(0, 2, 3, 1).reshape(M, N)
        else:
            y = a.float().view(M, K) @ w.float().t()
        if bias is not None:
            y = y + bias
        if rowvec is not None:
            y = y + rowvec.repeat_interleave(rows_per_sample, 0)
        if residual is not None:
            y = y + residual.float().view(M, N)
        if gn is not None:  # the contract of iir_gemm_args.gn_sums
            assert int(gn.abs().sum()) == 0, "accumulator not zero on entry"
            n_s, groups = gn.shape[0], gn.shape[1]
            assert rows_per_sample * n_s == M and ops.gn_eligible(N=N, groups=groups, rows_per_sample=rows_per_sample, conv=conv, residual=residual)
            t = y.double().view(n_s, M // n_s, groups, N // groups)
            gn[..., 0] += torch.round(t.sum((1, 3)) * S1).long()
            gn[..., 1] += torch.round((t * t).sum((1, 3)) * S2).long()
        out.copy_(y.to(out.dtype))
        return out

    def groupnorm(self, x, gamma, beta, out, *, n_img, HW, C, groups=32, eps=1e-5, silu=False, scratch_owner=None):
        self.calls.append("groupnorm")
        y = F.group_norm(x.float().view(n_img, HW, C).permute(0, 2, 1), groups, gamma, beta, eps).permute(0, 2, 1).reshape(n_img * HW, C)
        out.copy_((F.silu(y) if silu else y).to(out.dtype))
        return out

    def groupnorm_apply_sums(self, x, gamma, beta, sums, out, *, n_img, HW, C, groups=32, eps=1e-5, silu=False):
        self.calls.append("groupnorm_apply_sums")
        assert tuple(sums.shape) == (n_img, groups, 2) and sums.dtype == torch.int64
        count = HW * (C // groups)
        mean = sums[..., 0].double() / S1 / count
        var = (sums[..., 1].double() / S2 / count - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        xs = x.double().view(n_img, HW, groups, C // groups)
        y = ((xs - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(n_img * HW, C) * gamma.double() + beta.double()
        y = y.float()
        out.copy_((F.silu(y) if silu else y).to(out.dtype))
        return out

    def memset_zero(self, t):
        self.calls.append("memset")
        return t.zero_()

    def im2col3x3_s2(self, x, out, *, n_img, H, W, C, asym=False):
        self.calls.append("im2col")
        xp = F.pad(x.float().view(n_img, H, W, C).permute(0, 3, 1, 2), (1, 1, 1, 1))
        cols = F.unfold(xp, 3, stride=2)  # [n, C*9, L], channel-major
        L = cols.shape[-1]
        cols = cols.view(n_img, C, 9, L).permute(0, 3, 2, 1).reshape(n_img * L, 9 * C)  # tap-major like the kernel
        out.copy_(cols.to(out.dtype))
        return out

    def linear_small(self, x, w, bias, out, *, M, N, K, act=0):
        y = x.float() @ w.float().t()
        out.copy_(y + bias if bias is not None else y)
        return out


def _rt(fuse):
    rt = nn.Runtime.__new__(nn.Runtime)  # the real constructor refuses non-CUDA devices
    rt.device, rt.precision, rt.tc = torch.device("cpu"), "fp16", True
    rt.act_dtype = rt.w_dtype = torch.float16
    rt.lora_enabled = False
    rt.temb_bank, rt.adaln_bank = nn.SmallLinearBank(rt), nn.SmallLinearBank(rt)
    rt.gn_fuse, rt._gn_arena, rt._gn_used = fuse, None, 0
    return rt


class _Src:
    def __init__(self, seed=0):
        self.g = torch.Generator().manual_seed(seed)
        self.t = {}

    def has(self, k):
        return True

    def get(self, k):
        return self.t[k]

    def get_lora(self, m):
        return None

    def add(self, name, *shape, scale=None):
        t = torch.randn(*shape, generator=self.g)
        self.t[name] = t * (scale if scale is not None else (t[0].numel() ** -0.5 if t.ndim > 1 else 0.1))
        if name.endswith("norm1.weight") or name.endswith("norm2.weight"):
            self.t[name] = 1.0 + self.t[name]


def _resnet_src(src, p, cin, cout, T):
    src.add(p + ".norm1.weight", cin)
    src.add(p + ".norm1.bias", cin)
    src.add(p + ".conv1.weight", cout, cin, 3, 3)
    src.add(p + ".conv1.bias", cout)
    src.add(p + ".time_emb_proj.weight", cout, T)
    src.add(p + ".time_emb_proj.bias", cout)
    src.add(p + ".norm2.weight", cout)
    src.add(p + ".norm2.bias", cout)
    src.add(p + ".conv2.weight", cout, cout, 3, 3)
    src.add(p + ".conv2.bias", cout)
    if cin != cout:
        src.add(p + ".conv_shortcut.weight", cout, cin, 1, 1)
        src.add(p + ".conv_shortcut.bias", cout)


def _run(fuse, monkeypatch):
    fake = _Fake()
    for name in ("gemm", "groupnorm", "groupnorm_apply_sums", "memset_zero", "im2col3x3_s2", "linear_small"):
        monkeypatch.setattr(ops, name, getattr(fake, name))
    monkeypatch.setattr(ops, "cast2d", lambda x, ld_in, out, ld_out, rows, cols: out.copy_(x.view(rows, cols).to(out.dtype)))
    cfg = type("Cfg", (), {"norm_num_groups": 32, "norm_eps": 1e-5})()
    rt = _rt(fuse)
    src = _Src()
    T, C0, C1 = 64, 64, 128
    _resnet_src(src, "r0", C0, C0, T)
    _resnet_src(src, "r1", C0, C0, T)
    src.add("down.conv.weight", C0, C0, 3, 3)
    src.add("down.conv.bias", C0)
    _resnet_src(src, "r2", C0, C1, T)
    r0, r1 = nn.ResnetBlock2D(rt, src, "r0", cfg, C0, C0), nn.ResnetBlock2D(rt, src, "r1", cfg, C0, C0)
    down = nn.Downsample2D(rt, src, "down")
    r2 = nn.ResnetBlock2D(rt, src, "r2", cfg, C0, C1)
    n, H, W = 2, 16, 16
    x0 = torch.randn(n * H * W, C0, generator=torch.Generator().manual_seed(7))
    temb = torch.randn(n, T, generator=torch.Generator().manual_seed(8))
    outs = []
    for it in range(2):  # two forwards: the arena is cleared and re-used
        rt.new_forward()
        x = nn.FMap(x0.clone(), n, H, W, C0)
        x = r0(x, temb)
        used_after_r0 = rt._gn_used
        x = r1(x, temb)
        x = down(x, gn_groups=32)
        x = r2(x, temb)
        outs.append(x.t.clone())
    return fake, rt, outs, used_after_r0


def test_fused_groupnorm_plumbing_matches_the_two_kernel_path(monkeypatch):
    fake1, rt1, outs1, used = _run(True, monkeypatch)
    fake0, rt0, outs0, _ = _run(False, monkeypatch)
    # default path: no memset, no sums, 6 two-kernel GroupNorms per forward
    assert "memset" not in fake0.calls and "gemm+gn" not in fake0.calls and fake0.calls.count("groupnorm") == 12
    assert rt0._gn_arena is None
    # fused path: one memset per forward; r0.norm1 (input not produced by a GEMM) is the only fall-back; the other five
    # GroupNorms of a forward read the sums of conv1 / conv2 / the downsampler GEMM
    assert fake1.calls.count("memset") == 2
    assert fake1.calls.count("groupnorm") == 2 and fake1.calls.count("groupnorm_apply_sums") == 10
    # producers per forward: r0 (conv1, conv2), r1 (conv1, conv2), downsampler, r2 (conv1, conv2) = 7; the shortcut GEMM none
    assert fake1.calls.count("gemm+gn") == 14
    assert used == 2 * (2 * 32 * 2) and rt1._gn_used == 7 * (2 * 32 * 2)  # sites handed out in order, reset by new_forward
    for a, b in zip(outs1, outs0):
        assert float((a - b).norm() / b.norm()) < 2e-3  # the two paths normalise the same tensors
    assert torch.equal(outs1[0], outs1[1])               # second forward through the re-used arena: identical


def test_arena_exhaustion_and_ineligible_shapes_fall_back(monkeypatch):
    rt = _rt(True)
    monkeypatch.setattr(ops, "memset_zero", lambda t: t.zero_())
    monkeypatch.setattr(nn.Runtime, "GN_ARENA_WORDS", 300)
    assert rt.gn_site(2, 32) is None          # before the first forward there is no arena
    rt.new_forward()
    a, b = rt.gn_site(2, 32), rt.gn_site(2, 32)
    assert a.shape == (2, 32, 2) and b.data_ptr() == a.data_ptr() + 128 * 8
    assert rt.gn_site(2, 32) is None          # 300 words hold two sites of 128: the third caller falls back
    assert not ops.gn_eligible(N=320, groups=32, rows_per_sample=1000)                                   # rows of a warp would straddle samples
    assert not ops.gn_eligible(N=96, groups=32, rows_per_sample=1024)                                    # 3 channels per group: odd
    assert not ops.gn_eligible(N=320, groups=32, rows_per_sample=64, conv=dict(n_img=1, H=2, W=32, Cin=64))  # < 4 rows
    assert ops.gn_eligible(N=320, groups=32, rows_per_sample=64, conv=dict(n_img=1, H=8, W=8, Cin=64))


def test_the_gpu_child_script_itself_is_sound(monkeypatch):
    """tests/gn_fuse_child.py (the kernel checks the driver's GPU run will execute) against the torch emulations: a Python slip
    in the checker must not make a correct kernel look broken — reference computations, shapes, tolerances and the expected
    rejection of an ineligible launch are exercised here on the CPU."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("gn_fuse_child", os.path.join(os.path.dirname(os.path.abspath(__file__)), "gn_fuse_child.py"))
    child = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(child)
    fake = _Fake()

    def gemm(a, w, out, **kw):
        gn, conv = kw.get("gn"), kw.get("conv")
        if gn is not None and not ops.gn_eligible(N=kw["N"], groups=gn.shape[1], rows_per_sample=kw.get("rows_per_sample", 0), conv=conv,
                                                  residual=kw.get("residual")):
            raise ops._lib.IIRError("iir_gemm_tc failed (-1): gn_sums needs rows_per_sample % 32 == 0")  # what the C side answers
        return fake.gemm(a, w, out, **kw)

    monkeypatch.setattr(child, "DEV", "cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    for name in ("groupnorm", "groupnorm_apply_sums", "memset_zero"):
        monkeypatch.setattr(ops, name, getattr(fake, name))
    monkeypatch.setattr(ops, "gemm", gemm)
    checks = {}
    child.run_kernels(checks)
    assert any(k.endswith(":rejected") for k in checks)
    sums = [v for k, v in checks.items() if k.endswith(":sumsq_rel_err")]
    assert len(sums) >= 12 and max(sums) < 1e-5          # the emulation is exact up to fixed-point rounding
    assert max(v for k, v in checks.items() if "_vs_two_kernel" in k) < 2e-3
