"""Oracle schedulers (CPU fp32): LCM single-step previewer and DDPM ancestral step.

* ``LCMSingleStepScheduler`` restates schedulers/lcm_single_step_scheduler.py:193-249,401-489,492-513
  (pinned by tests/golden/lcm_scheduler.pt, produced by running that file verbatim).
* ``DDPMScheduler`` restates diffusers==0.28.1 ``DDPMScheduler`` as configured by SDXL's
  scheduler_config.json.  diffusers is absent here: **parity unpinned**,
  checked by closed-form tests only.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def _scaled_linear_alphas_cumprod(num_train_timesteps, beta_start, beta_end):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return betas, torch.cumprod(1.0 - betas, dim=0)


class LCMSingleStepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, timestep_scaling=10.0,
                 original_inference_steps=50):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, timestep_scaling=timestep_scaling,
                                      prediction_type="epsilon", clip_sample=False,
                                      original_inference_steps=original_inference_steps)
        self.betas, self.alphas_cumprod = _scaled_linear_alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.init_noise_sigma = 1.0
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def get_scalings_for_boundary_condition_discrete(self, timestep):
        sigma_data = 0.5
        scaled = timestep * self.config.timestep_scaling
        c_skip = sigma_data ** 2 / (scaled ** 2 + sigma_data ** 2)
        c_out = scaled / (scaled ** 2 + sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor(timestep, dtype=torch.int64)
        if timestep.ndim == 0:
            timestep = timestep.unsqueeze(0)
        shape = (timestep.shape[0],) + (1,) * (sample.ndim - 1)
        alpha_prod_t = self.alphas_cumprod.gather(-1, timestep).reshape(shape)
        beta_prod_t = 1 - alpha_prod_t
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(timestep)
        c_skip, c_out = c_skip.reshape(shape), c_out.reshape(shape)
        x0 = (sample - torch.sqrt(beta_prod_t) * model_output) / torch.sqrt(alpha_prod_t)
        denoised = c_out * x0 + c_skip * sample
        return SimpleNamespace(denoised=denoised) if return_dict else (denoised,)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)
        sa = (ac[timesteps] ** 0.5).flatten()
        sb = ((1 - ac[timesteps]) ** 0.5).flatten()
        while sa.ndim < original_samples.ndim:
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class DDPMScheduler:
    """leading spacing, steps_offset=1, fixed_small variance, epsilon prediction, no clipping."""

    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      beta_start=beta_start, beta_end=beta_end)
        self.betas, self.alphas_cumprod = _scaled_linear_alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.one = torch.tensor(1.0)
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            self.custom_timesteps = True
            ts = np.array(timesteps, dtype=np.int64)
        else:
            self.custom_timesteps = False
            self.num_inference_steps = num_inference_steps
            ratio = self.config.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def previous_timestep(self, timestep):
        if self.custom_timesteps:
            idx = (self.timesteps == timestep).nonzero(as_tuple=True)[0][0]
            return torch.tensor(-1) if idx == self.timesteps.shape[0] - 1 else self.timesteps[idx + 1]
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def coefficients(self, timestep):
        """(alpha_prod_t, c_x0, c_xt, sigma) for one step; sigma = 0 when t == 0."""
        t = int(timestep)
        prev_t = int(self.previous_timestep(timestep))
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

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, noise=None):
        a_t, c_x0, c_xt, sigma = self.coefficients(timestep)
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        prev = c_x0 * x0 + c_xt * sample
        if int(timestep) > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + sigma * noise
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)
