"""Schedulers with the reference's API, stepping on the fused sm_100a kernels.

``LCMSingleStepScheduler`` — schedulers/lcm_single_step_scheduler.py (``from_config``, ``step``,
``set_timesteps``-less single-step use, ``add_noise``, ``scale_model_input``).
``DDPMScheduler`` — diffusers==0.28.1 DDPMScheduler as configured by SDXL:
leading spacing, steps_offset 1, fixed_small variance, epsilon prediction, no clipping.
The per-timestep scalar coefficients are computed on the host in fp32 exactly like the reference's
0-dim tensor arithmetic; the elementwise work is one kernel launch per step.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import ops


def _alphas_cumprod(num_train_timesteps, beta_start, beta_end):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return betas, torch.cumprod(1.0 - betas, dim=0)


def _randn(shape, generator, device):
    """diffusers randn_tensor: a CPU generator draws on the CPU (bit-identical to the oracle) and the
    sample is then moved to the device."""
    gdev = generator.device.type if generator is not None else torch.device(device).type
    if gdev == "cpu":
        return torch.randn(shape, generator=generator, dtype=torch.float32).to(device)
    return torch.randn(shape, generator=generator, dtype=torch.float32, device=device)


def _add_noise(alphas_cumprod, original_samples, noise, timesteps):
    """sqrt(abar_t) x0 + sqrt(1 - abar_t) noise with abar indexed PER SAMPLE (lcm_single_step_scheduler.py:492-513; same
    code in diffusers' DDPMScheduler): one launch when every sample shares the timestep (the pipeline's case,
    pipelines/sdxl_instantir.py:931-939), one per sample otherwise."""
    ts = torch.as_tensor(timesteps).reshape(-1).tolist()
    x0, z = original_samples.float().contiguous(), noise.float().contiguous()
    B = x0.shape[0]
    if len(ts) == 1:
        ts = ts * B
    if len(ts) != B:
        raise ValueError(f"add_noise: {len(ts)} timesteps for a batch of {B}")
    out = torch.empty_like(x0)
    if all(t == ts[0] for t in ts):
        ops.add_noise(x0, z, out, alpha_prod_t=float(alphas_cumprod[int(ts[0])]))
    else:
        for b, t in enumerate(ts):
            ops.add_noise(x0[b], z[b], out[b], alpha_prod_t=float(alphas_cumprod[int(t)]))
    return out


class LCMSingleStepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 timestep_scaling=10.0, prediction_type="epsilon", clip_sample=False, **unused):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or clip_sample:
            raise NotImplementedError("only the SDXL configuration (scaled_linear, epsilon, no clipping) is built")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, timestep_scaling=timestep_scaling,
                                      prediction_type=prediction_type, clip_sample=clip_sample)
        self.betas, self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.init_noise_sigma = 1.0
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kw):
        cfg = dict(config if isinstance(config, dict) else vars(config))
        cfg.update(kw)
        return cls(**cfg)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def get_scalings_for_boundary_condition_discrete(self, timestep):
        sigma_data = 0.5
        scaled = torch.as_tensor(timestep) * self.config.timestep_scaling
        c_skip = sigma_data ** 2 / (scaled ** 2 + sigma_data ** 2)
        c_out = scaled / (scaled ** 2 + sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, out=None):
        """denoised = c_out * (sample - sqrt(1-abar) eps)/sqrt(abar) + c_skip * sample, fp32 result."""
        t = int(timestep)
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(torch.tensor(t, dtype=torch.int64))
        x = sample if sample.dtype == torch.float32 else sample.float()
        x = x.contiguous()
        if out is None:
            out = torch.empty_like(x)
        ops.lcm_step(model_output.contiguous(), x, out, alpha_prod_t=float(self.alphas_cumprod[t]),
                     c_skip=float(c_skip), c_out=float(c_out))
        return SimpleNamespace(denoised=out) if return_dict else (out,)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)


class DDPMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 steps_offset=1, timestep_spacing="leading", variance_type="fixed_small", prediction_type="epsilon",
                 clip_sample=False, **unused):
        if (beta_schedule, timestep_spacing, variance_type, prediction_type, clip_sample) != (
                "scaled_linear", "leading", "fixed_small", "epsilon", False):
            raise NotImplementedError("only SDXL's scheduler_config.json settings are built (SURVEY Appendix C.4)")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      steps_offset=steps_offset, timestep_spacing=timestep_spacing,
                                      prediction_type=prediction_type, beta_schedule=beta_schedule)
        self.betas, self.alphas_cumprod = _alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.one = torch.tensor(1.0)
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    @classmethod
    def from_config(cls, config, **kw):
        cfg = dict(config if isinstance(config, dict) else vars(config))
        cfg.update(kw)
        return cls(**cfg)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if num_inference_steps is not None and timesteps is not None:
            raise ValueError("Can only pass one of `num_inference_steps` or `custom_timesteps`.")
        if timesteps is not None:
            ts = np.array(timesteps, dtype=np.int64)
            if (ts[1:] >= ts[:-1]).any():
                raise ValueError("`custom_timesteps` must be in descending order.")
            if ts[0] >= self.config.num_train_timesteps:
                raise ValueError(f"`timesteps` must start before `self.config.train_timesteps`: {self.config.num_train_timesteps}.")
            self.custom_timesteps = True
        else:
            if num_inference_steps > self.config.num_train_timesteps:
                raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
            self.custom_timesteps = False
            self.num_inference_steps = num_inference_steps
            ratio = self.config.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def previous_timestep(self, timestep):
        if self.custom_timesteps:
            idx = (self.timesteps == int(timestep)).nonzero(as_tuple=True)[0][0]
            return -1 if idx == self.timesteps.shape[0] - 1 else int(self.timesteps[idx + 1])
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return int(timestep) - self.config.num_train_timesteps // n

    def coefficients(self, timestep):
        """(abar_t, c_x0, c_xt, sigma) as fp32 0-dim tensors, sigma = 0 at t == 0."""
        t, prev_t = int(timestep), self.previous_timestep(timestep)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c_x0 = (a_prev ** 0.5 * cur_beta) / b_t
        c_xt = cur_alpha ** 0.5 * b_prev / b_t
        var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return a_t, c_x0, c_xt, sigma

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, *, guidance=None, noise=None):
        """x_{t-1} from eps.  Extension: when `model_output` holds the CFG pair [uncond; cond] (2x the
        batch of `sample`) and `guidance` is given, the CFG combine (pipelines/sdxl_instantir.py:1619-1621)
        runs inside the same kernel."""
        a_t, c_x0, c_xt, sigma = self.coefficients(timestep)
        x = sample.contiguous()
        if noise is None and int(timestep) > 0:
            noise = _randn(tuple(sample.shape), generator, sample.device)
        eps = model_output.contiguous()
        if guidance is not None:
            if eps.shape[0] != 2 * x.shape[0]:
                raise ValueError("guidance= needs model_output = cat([uncond, cond])")
            e_u, e_c = eps[: x.shape[0]], eps[x.shape[0]:]
        else:
            e_u, e_c, guidance = eps, None, 1.0
        prev, x0 = torch.empty_like(x), torch.empty_like(x)
        ops.cfg_ddpm_step(e_u, e_c, x, noise, prev, x0, guidance=float(guidance), alpha_prod_t=float(a_t),
                          c_x0=float(c_x0), c_xt=float(c_xt), sigma=float(sigma))
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        return _add_noise(self.alphas_cumprod, original_samples, noise, timesteps)
