"""VAE encode / decode on the sm_100a kernels (the first "next" row): the once-per-image
``vae.encode(image).latent_dist.sample() * scaling_factor`` of ``pipelines/sdxl_instantir.py:1370-1376`` and
``vae.decode(latents / scaling_factor)`` of ``:1670-1704``.

``AutoencoderKL.decode(z, return_dict)`` / ``.config.scaling_factor`` / ``.config.force_upcast`` mirror
``module/diffusers_vae/autoencoder_kl.py:270-300``; the decoder follows ``module/diffusers_vae/vae.py:185-350``
(conv_in -> UNetMidBlock2D -> UpDecoderBlock2D x4 -> GroupNorm -> SiLU -> conv_out).  State-dict keys are the
checkpoint's (``post_quant_conv.*``, ``decoder.*``).

Kernel plan (all existing kernels but one): 3x3 convs = the tcgen05 implicit GEMM with the residual add in its
epilogue; GroupNorm(+SiLU) = gn_stats/gn_apply; nearest-2x + conv = upsample2x + implicit GEMM; conv_in (4 -> C) and
conv_out (C -> 3, padded to 4) = the direct small-channel kernels.  The mid-block attention has ONE head of dim C
(512): S = Q Kᵀ and O = P V run on the GEMM kernel in query chunks, with ``iir_softmax_rows`` between them; V is
produced transposed by a GEMM (Vᵀ = W_v Xᵀ) and its bias is added after P V (rows of P sum to 1).

The reference runs the VAE in fp32 (``force_upcast``, fp16 overflows in it); here ``bf16`` operands have the range,
and ``fp32`` precision is the check mode.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .nn import Conv3x3, FMap, GroupNorm, Linear, Runtime, _ShortcutSrc, _bias


class VaeConfig(SimpleNamespace):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.13025, force_upcast=True, **unused):
        super().__init__(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                         block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                         norm_num_groups=norm_num_groups, scaling_factor=scaling_factor, force_upcast=force_upcast)


def vae_param_shapes(cfg) -> dict:
    """{state-dict key: shape} of the whole AutoencoderKL (encoder, quant_conv, post_quant_conv, decoder)."""
    d = dict(vae_decoder_param_shapes(cfg))
    ch, L = cfg.block_out_channels, cfg.latent_channels

    def conv(p, co, ci, k):
        d[p + ".weight"], d[p + ".bias"] = (co, ci, k, k), (co,)

    def norm(p, c):
        d[p + ".weight"], d[p + ".bias"] = (c,), (c,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    out = ch[0]
    for i in range(len(ch)):
        prev, out = out, ch[i]
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else out, out)
        if i != len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    resnet("encoder.mid_block.resnets.0", ch[-1], ch[-1])
    resnet("encoder.mid_block.resnets.1", ch[-1], ch[-1])
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch[-1])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        d[f"{a}.{n}.weight"], d[f"{a}.{n}.bias"] = (ch[-1], ch[-1]), (ch[-1],)
    norm("encoder.conv_norm_out", ch[-1])
    conv("encoder.conv_out", 2 * L, ch[-1], 3)
    conv("quant_conv", 2 * L, 2 * L, 1)
    return d


def vae_decoder_param_shapes(cfg) -> dict:
    """{state-dict key: shape} of post_quant_conv + decoder (for weights.RandomSource at full size)."""
    d = {}
    ch, L = cfg.block_out_channels, cfg.latent_channels

    def conv(p, co, ci, k):
        d[p + ".weight"], d[p + ".bias"] = (co, ci, k, k), (co,)

    def norm(p, c):
        d[p + ".weight"], d[p + ".bias"] = (c,), (c,)

    def lin(p, co, ci):
        d[p + ".weight"], d[p + ".bias"] = (co, ci), (co,)

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    conv("post_quant_conv", L, L, 1)
    conv("decoder.conv_in", ch[-1], L, 3)
    resnet("decoder.mid_block.resnets.0", ch[-1], ch[-1])
    resnet("decoder.mid_block.resnets.1", ch[-1], ch[-1])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch[-1])
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", ch[-1], ch[-1])
    rev = list(reversed(ch))
    out = rev[0]
    for i in range(len(ch)):
        prev, out = out, rev[i]
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out, out)
        if i != len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", out, out, 3)
    norm("decoder.conv_norm_out", ch[0])
    conv("decoder.conv_out", cfg.out_channels, ch[0], 3)
    return d


class _Resnet:
    """diffusers ResnetBlock2D without a time embedding (eps 1e-6): GN+SiLU -> conv1 -> GN+SiLU -> conv2 with the
    residual (input, or its 1x1 shortcut GEMM) added in conv2's epilogue."""

    def __init__(self, rt, src, p, c_in, c_out, groups):
        self.rt = rt
        self.norm1 = GroupNorm(rt, src, p + ".norm1", c_in, groups, 1e-6)
        self.conv1 = Conv3x3(rt, src, p + ".conv1")
        self.norm2 = GroupNorm(rt, src, p + ".norm2", c_out, groups, 1e-6)
        self.conv2 = Conv3x3(rt, src, p + ".conv2")
        self.conv_shortcut = Linear(rt, _ShortcutSrc(src), p + ".conv_shortcut") if c_in != c_out else None

    def __call__(self, x: FMap) -> FMap:
        rt = self.rt
        h = self.conv1(self.norm1(x, silu=True))
        h = self.norm2(h, silu=True)
        res = x.t
        if self.conv_shortcut is not None:
            a = x.t
            if a.dtype != rt.act_dtype:
                a = rt.empty(x.M, x.C)
                ops.cast2d(x.t, x.C, a, x.C, rows=x.M, cols=x.C)
            res = self.conv_shortcut(a, x.M, out_dtype=torch.float32)
        return self.conv2(h, out_dtype=torch.float32, residual=res)


class _MidAttention:
    """diffusers Attention of the VAE mid block driven by AttnProcessor2_0 (the reference's copy:
    module/ip_adapter/attention_processor.py:337-414 with a 4-D input): GroupNorm(eps 1e-6) -> q/k/v (bias) ->
    softmax(q kᵀ / sqrt(C)) v, ONE head -> to_out (bias) -> + input."""

    Q_CHUNK = 8192  # query rows per S = Q Kᵀ launch: bounds the fp32 score buffer (8192 x HW) at any image size

    def __init__(self, rt, src, p, C, groups):
        self.rt, self.C = rt, C
        self.group_norm = GroupNorm(rt, src, p + ".group_norm", C, groups, 1e-6)
        self.to_q = Linear(rt, src, p + ".to_q")
        self.to_k = Linear(rt, src, p + ".to_k")
        self.w_v = src.get(p + ".to_v.weight").to(rt.w_dtype).contiguous()  # A operand of Vᵀ = W_v Xᵀ
        self.b_v = _bias(src, p + ".to_v")
        self.to_out = Linear(rt, src, p + ".to_out.0")

    def __call__(self, x: FMap) -> FMap:
        rt, C = self.rt, self.C
        HW = x.H * x.W
        xn = self.group_norm(x, silu=False)
        out = rt.stream(x.M, C)
        rows_max = min(self.Q_CHUNK, HW)
        s_buf = torch.empty(rows_max, HW, device=rt.device, dtype=torch.float32)
        p_buf = rt.empty(rows_max, HW)
        for b in range(x.n):
            xb = xn.t[b * HW:(b + 1) * HW]
            q, k = self.to_q(xb, HW), self.to_k(xb, HW)
            vt = rt.empty(C, HW)
            ops.gemm(self.w_v, xb, vt, M=C, N=HW, K=C, tc=rt.tc)
            o = rt.empty(HW, C)
            for r0 in range(0, HW, rows_max):
                rows = min(rows_max, HW - r0)
                ops.gemm(q[r0:r0 + rows], k, s_buf[:rows], M=rows, N=HW, K=C, tc=rt.tc)
                ops.softmax_rows(s_buf[:rows], p_buf[:rows], scale=C ** -0.5)
                ops.gemm(p_buf[:rows], vt, o[r0:r0 + rows], M=rows, N=C, K=HW, bias=self.b_v, tc=rt.tc)
            sl = slice(b * HW, (b + 1) * HW)
            self.to_out(o, HW, out=out[sl], residual=x.t[sl])
        return FMap(out, x.n, x.H, x.W, C)


class Encoder:
    """module/diffusers_vae/vae.py:46-182: conv_in -> DownEncoderBlock2D x4 (resnets, then a stride-2 conv padded
    bottom/right only) -> UNetMidBlock2D -> GroupNorm -> SiLU -> conv_out (2 x latent channels)."""

    def __init__(self, rt: Runtime, src, cfg, p="encoder"):
        self.rt, self.cfg = rt, cfg
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in_w = src.get(p + ".conv_in.weight").permute(0, 2, 3, 1).contiguous().float()
        self.conv_in_b = src.get(p + ".conv_in.bias").contiguous().float()
        self.down_blocks = []
        out = ch[0]
        for i in range(len(ch)):
            prev, out = out, ch[i]
            resnets = [_Resnet(rt, src, f"{p}.down_blocks.{i}.resnets.{j}", prev if j == 0 else out, out, g)
                       for j in range(cfg.layers_per_block)]
            down = Conv3x3(rt, src, f"{p}.down_blocks.{i}.downsamplers.0.conv", stride=2, asym=True) if i != len(ch) - 1 else None
            self.down_blocks.append((resnets, down))
        self.mid = [_Resnet(rt, src, p + ".mid_block.resnets.0", ch[-1], ch[-1], g),
                    _MidAttention(rt, src, p + ".mid_block.attentions.0", ch[-1], g),
                    _Resnet(rt, src, p + ".mid_block.resnets.1", ch[-1], ch[-1], g)]
        self.conv_norm_out = GroupNorm(rt, src, p + ".conv_norm_out", ch[-1], g, 1e-6)
        self.conv_out = Conv3x3(rt, src, p + ".conv_out")

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, 3, H, W] fp32 in [-1, 1] -> [B, 2*latent, H/8, W/8] fp32 (NCHW view of NHWC memory)."""
        rt, cfg = self.rt, self.cfg
        B, Ci, H, W = x.shape
        n_down = len(cfg.block_out_channels) - 1
        if H % (1 << n_down) or W % (1 << n_down):
            raise ValueError(f"image size {H}x{W} must be a multiple of {1 << n_down}")
        c0 = cfg.block_out_channels[0]
        h = rt.stream(B * H * W, c0)
        ops.conv3x3_direct(x.contiguous(), self.conv_in_w, self.conv_in_b, h, in_nchw=True, out_nchw=False,
                           n_img=B, H=H, W=W, Cin=Ci, Cout=c0)
        h = FMap(h, B, H, W, c0)
        for resnets, down in self.down_blocks:
            for r in resnets:
                h = r(h)
            if down is not None:
                h = down(h, out_dtype=torch.float32)
        for m in self.mid:
            h = m(h)
        y = self.conv_norm_out(h, silu=True)
        return self.conv_out(y, out_dtype=torch.float32).nchw()


class Decoder:
    def __init__(self, rt: Runtime, src, cfg, p="decoder"):
        self.rt, self.cfg = rt, cfg
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in_w = src.get(p + ".conv_in.weight").permute(0, 2, 3, 1).contiguous().float()
        self.conv_in_b = src.get(p + ".conv_in.bias").contiguous().float()
        self.mid = [_Resnet(rt, src, p + ".mid_block.resnets.0", ch[-1], ch[-1], g),
                    _MidAttention(rt, src, p + ".mid_block.attentions.0", ch[-1], g),
                    _Resnet(rt, src, p + ".mid_block.resnets.1", ch[-1], ch[-1], g)]
        rev = list(reversed(ch))
        self.up_blocks = []
        out = rev[0]
        for i in range(len(ch)):
            prev, out = out, rev[i]
            resnets = [_Resnet(rt, src, f"{p}.up_blocks.{i}.resnets.{j}", prev if j == 0 else out, out, g)
                       for j in range(cfg.layers_per_block + 1)]
            up = Conv3x3(rt, src, f"{p}.up_blocks.{i}.upsamplers.0.conv") if i != len(ch) - 1 else None
            self.up_blocks.append((resnets, up))
        self.conv_norm_out = GroupNorm(rt, src, p + ".conv_norm_out", ch[0], g, 1e-6)
        # conv_out: C -> 3, padded with a zero output channel so that the direct small-C_out kernel (C_out == 4) applies
        w = src.get(p + ".conv_out.weight").permute(0, 2, 3, 1).contiguous().float()
        b = src.get(p + ".conv_out.bias").float()
        self.n_out = w.shape[0]
        if self.n_out > 4:
            raise NotImplementedError("VAE decoders with more than 4 output channels")
        self.conv_out_w = torch.zeros(4, 3, 3, ch[0], device=w.device)
        self.conv_out_w[:self.n_out] = w
        self.conv_out_b = torch.zeros(4, device=w.device)
        self.conv_out_b[:self.n_out] = b

    def __call__(self, z: torch.Tensor) -> torch.Tensor:
        """z [B, latent, h, w] fp32 -> image [B, 3, 8h, 8w] fp32 (NCHW)."""
        rt, cfg = self.rt, self.cfg
        B, L, h, w = z.shape
        c_mid = cfg.block_out_channels[-1]
        x = rt.stream(B * h * w, c_mid)
        ops.conv3x3_direct(z.contiguous(), self.conv_in_w, self.conv_in_b, x, in_nchw=True, out_nchw=False,
                           n_img=B, H=h, W=w, Cin=L, Cout=c_mid)
        x = FMap(x, B, h, w, c_mid)
        for m in self.mid:
            x = m(x)
        for resnets, up in self.up_blocks:
            for r in resnets:
                x = r(x)
            if up is not None:
                x = up(x, out_dtype=torch.float32, up2=True)
        y = self.conv_norm_out(x, silu=True)
        img = torch.empty(B, 4, x.H, x.W, device=rt.device, dtype=torch.float32)
        ops.conv3x3_direct(y.t, self.conv_out_w, self.conv_out_b, img, in_nchw=False, out_nchw=True,
                           n_img=B, H=x.H, W=x.W, Cin=x.C, Cout=4)
        return img[:, :self.n_out]


class DiagonalGaussianDistribution:
    """module/diffusers_vae/vae.py DiagonalGaussianDistribution over encoder moments [B, 2L, h, w] (mean | logvar):
    ``sample(generator)`` = mean + exp(0.5 clamp(logvar, -30, 20)) * noise, ``mode()`` = mean."""

    def __init__(self, moments: torch.Tensor):
        self.parameters = moments.contiguous()
        self.shape = (moments.shape[0], moments.shape[1] // 2) + tuple(moments.shape[2:])

    def sample(self, generator=None, scale: float = 1.0) -> torch.Tensor:
        from .schedulers import _randn

        noise = _randn(self.shape, generator, self.parameters.device)
        return ops.gaussian_sample(self.parameters, noise, torch.empty(self.shape, device=self.parameters.device), scale=scale)

    def mode(self) -> torch.Tensor:
        return ops.gaussian_sample(self.parameters, None, torch.empty(self.shape, device=self.parameters.device))


class AutoencoderKL:
    """the reference's AutoencoderKL as the pipeline uses it: ``encode(x).latent_dist.sample()`` and ``decode(z)``.
    A weight source without ``encoder.*`` keys gives a decode-only model."""

    def __init__(self, cfg, source, device="cuda", precision="bf16"):
        self.config = cfg if isinstance(cfg, VaeConfig) else VaeConfig(**(cfg if isinstance(cfg, dict) else cfg.to_dict()))
        self.rt = Runtime(device, precision)
        L = self.config.latent_channels
        # post_quant_conv (1x1, L -> L) as a 3x3 direct conv whose only non-zero tap is the centre: exact
        w = source.get("post_quant_conv.weight").float().reshape(L, L)
        self.pq_w = torch.zeros(L, 3, 3, L, device=w.device)
        self.pq_w[:, 1, 1, :] = w
        self.pq_b = source.get("post_quant_conv.bias").contiguous().float()
        self.decoder = Decoder(self.rt, source, self.config)
        self.dtype = self.rt.act_dtype
        self.encoder = None
        if source.has("encoder.conv_in.weight"):
            self.encoder = Encoder(self.rt, source, self.config)
            w = source.get("quant_conv.weight").float().reshape(2 * L, 2 * L)
            self.q_w = torch.zeros(2 * L, 3, 3, 2 * L, device=w.device)
            self.q_w[:, 1, 1, :] = w
            self.q_b = source.get("quant_conv.bias").contiguous().float()

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """module/diffusers_vae/autoencoder_kl.py:236-268: encoder -> quant_conv -> DiagonalGaussianDistribution."""
        if self.encoder is None:
            raise ValueError("this AutoencoderKL was built from a weight source without 'encoder.*' tensors")
        if x.ndim != 4 or x.shape[1] != self.config.in_channels:
            raise ValueError(f"expected images [B, {self.config.in_channels}, H, W], got {tuple(x.shape)}")
        x = x.to(device=self.rt.device, dtype=torch.float32).contiguous()
        h = self.encoder(x).contiguous()  # NHWC memory -> NCHW for the 1x1 quant_conv
        B, C2, hh, ww = h.shape
        moments = torch.empty(B, C2, hh, ww, device=self.rt.device, dtype=torch.float32)
        ops.conv3x3_direct(h, self.q_w, self.q_b, moments, in_nchw=True, out_nchw=True, n_img=B, H=hh, W=ww, Cin=C2, Cout=C2)
        dist = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (dist,)
        return SimpleNamespace(latent_dist=dist)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """module/diffusers_vae/autoencoder_kl.py:270-300: post_quant_conv then the decoder."""
        if z.ndim != 4 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"expected latents [B, {self.config.latent_channels}, h, w], got {tuple(z.shape)}")
        z = z.to(device=self.rt.device, dtype=torch.float32).contiguous()
        B, L, h, w = z.shape
        zq = torch.empty_like(z)
        ops.conv3x3_direct(z, self.pq_w, self.pq_b, zq, in_nchw=True, out_nchw=True, n_img=B, H=h, W=w, Cin=L, Cout=L)
        img = self.decoder(zq)
        if not return_dict:
            return (img,)
        return SimpleNamespace(sample=img)


def postprocess(image: torch.Tensor, output_type: str = "pt"):
    """VaeImageProcessor.postprocess (diffusers; call site pipelines/sdxl_instantir.py:1704): denormalise to [0, 1];
    'pt' -> tensor [B,3,H,W], 'np' -> float32 array [B,H,W,3]."""
    img = (image / 2 + 0.5).clamp(0, 1)
    if output_type == "pt":
        return img
    if output_type == "np":
        return img.permute(0, 2, 3, 1).float().cpu().numpy()
    raise NotImplementedError(f"output_type={output_type!r}: 'latent', 'pt' and 'np' are built (PIL conversion is host-side glue)")
