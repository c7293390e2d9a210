"""Helpers of the CPU loop tests (test_pipeline_loop_cpu.py, test_pipeline_parallel_cpu.py): the ORACLE UNet / Aggregator behind the
product objects' call surface, and torch emulations of the ops the loop and the schedulers launch themselves."""
from types import SimpleNamespace

import torch

from instantir_b200 import ops
from oracle import pipeline as opipe


class NullStream:
    """stand-in for torch.cuda.Stream on a box without a GPU: work runs in program order"""

    def __init__(self, device=None):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class OracleUNet:
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


class OracleAgg:
    weights_version = 0

    def __init__(self, o):
        self.o = o

    def __call__(self, sample, t_dev, **kw):
        t = torch.tensor(int(t_dev.reshape(-1)[0]), dtype=torch.int64)
        return self.o(sample, t, **kw)


def emulate_ops(setattr_fn):
    """install torch emulations of the ops the loop and the schedulers launch themselves; setattr_fn(obj, name, value)"""
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
        setattr_fn(ops, name, fn)
    setattr_fn(torch.cuda, "Stream", NullStream)
