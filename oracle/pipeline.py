"""Oracle denoising loop — restates pipelines/sdxl_instantir.py:1385-1666 (CPU fp32).

Inputs are already-encoded tensors (text/pooled/DINOv2 embeddings, LQ latent): the once-per-image
encoders are outside the hot path.  Quirks of the reference loop are kept:
the `(cond_scale>0.1).sum().item() > 0` gate (:1542), stale residuals multiplied by zero after
control ends (:1602-1603), the duplicate temb computation (:1516-1529).
"""
from __future__ import annotations

from typing import List, Optional

import torch


def step_masks(n_steps, preview_start, preview_end, control_guidance_start, control_guidance_end):
    """controlnet_keep / previewing — pipelines/sdxl_instantir.py:1415-1421."""
    keep, prev = [], []
    for i in range(n_steps):
        keep.append(1.0 - float(i / n_steps < control_guidance_start or (i + 1) / n_steps > control_guidance_end))
        prev.append(1.0 - float(i / n_steps < preview_start or (i + 1) / n_steps > preview_end))
    return keep, prev


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """pipelines/sdxl_instantir.py:181-192: pull the guided prediction's per-sample std back to the text branch's and
    blend by `guidance_rescale` (torch.std = unbiased)."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text, std_cfg = noise_pred_text.std(dim=dims, keepdim=True), noise_cfg.std(dim=dims, keepdim=True)
    noise_pred_rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * noise_pred_rescaled + (1 - guidance_rescale) * noise_cfg


@torch.no_grad()
def restore_latents(unet, aggregator, scheduler, previewer_scheduler, *, image, prompt_embeds,
                    negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds,
                    ip_image_embeds, add_time_ids, num_inference_steps=30, guidance_scale=7.0,
                    preview_start=0.0, preview_end=1.0, control_guidance_start=0.0, control_guidance_end=1.0,
                    controlnet_conditioning_scale=1.0, generator=None, init_latents_with_lq=True,
                    timesteps=None, record: Optional[dict] = None, guidance_rescale: float = 0.0,
                    max_steps: Optional[int] = None, adastep_restore: bool = False, denoising_end: Optional[float] = None,
                    reference_latents=None):
    """Returns the final latents [B,4,h,w]; `record` (if given) collects per-step tensors.

    image: LQ latent [B,4,h,w] (the reference accepts 4-channel tensors as latents, :1370-1382).
    ip_image_embeds: [2,B,S,D] stacked (negative, positive) DINOv2 tokens (:700-706).
    """
    do_cfg = guidance_scale > 1.0
    B = image.shape[0]
    scheduler.set_timesteps(num_inference_steps, timesteps=timesteps)
    ts = scheduler.timesteps
    n = len(ts)
    if init_latents_with_lq:  # init_latents (:931-939): noise drawn first from the user generator
        noise = torch.randn(image.shape, generator=generator, dtype=image.dtype)
        latents = scheduler_add_noise(scheduler, image, noise, ts[0])
    else:
        latents = torch.randn(image.shape, generator=generator, dtype=image.dtype) * scheduler.init_noise_sigma
    keep, previewing = step_masks(n, preview_start, preview_end, control_guidance_start, control_guidance_end)
    scales = controlnet_conditioning_scale if isinstance(controlnet_conditioning_scale, list) else [controlnet_conditioning_scale] * n
    if denoising_end is not None and isinstance(denoising_end, float) and 0 < denoising_end < 1:  # 8.1, :1469-1484
        n_train = scheduler.config.num_train_timesteps
        cutoff = int(round(n_train - denoising_end * n_train))
        ts = ts[: len([t for t in ts.tolist() if t >= cutoff])]
    add_text_embeds = pooled_prompt_embeds
    time_ids = add_time_ids
    if do_cfg:
        prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
        add_text_embeds = torch.cat([negative_pooled_prompt_embeds, add_text_embeds], dim=0)
        time_ids = torch.cat([time_ids, time_ids], dim=0)
        image = torch.cat([image] * 2, dim=0)
        image_embeds = [torch.cat([ip_image_embeds[0], ip_image_embeds[1]], dim=0).unsqueeze(1)]
    else:
        image_embeds = [ip_image_embeds[1].unsqueeze(1)]
    preview_factor = torch.ones((latents.shape[0], 1, 1, 1), dtype=latents.dtype)
    previewer_mean = torch.zeros_like(latents)  # :1488
    down_res = mid_res = None
    last_preview = None  # the reference's `preview_latent` is a plain local: it keeps the last gated step's value
    for i, t in enumerate(ts):
        if max_steps is not None and i >= max_steps:  # tests: only the first steps of a long schedule
            break
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        x_in = scheduler.scale_model_input(x_in, t)
        added = {"text_embeds": add_text_embeds, "time_ids": time_ids, "image_embeds": image_embeds}
        emb = unet.time_embedding(unet.get_time_embed(sample=x_in, timestep=t))
        emb = emb + unet.get_aug_embed(emb=emb, encoder_hidden_states=prompt_embeds, added_cond_kwargs=added)
        ca_kwargs = {"temb": emb}
        cond_scale = preview_factor.clamp(0.0, scales[i]) * keep[i]
        cond_scale = torch.cat([cond_scale] * 2) if do_cfg else cond_scale
        preview_latent = None
        if (cond_scale > 0.1).sum().item() > 0:
            if previewing[i] > 0:
                unet.enable_adapters()
                preview_noise = unet(x_in, t, encoder_hidden_states=prompt_embeds, cross_attention_kwargs=ca_kwargs,
                                     added_cond_kwargs=added, return_dict=False)[0]
                preview_latent = previewer_scheduler.step(preview_noise, t.to(dtype=torch.int64), x_in,
                                                          return_dict=False)[0]
                unet.disable_adapters()
            elif reference_latents is not None:  # :1579-1580
                preview_latent = torch.cat([reference_latents] * 2) if do_cfg else reference_latents
            else:
                preview_latent = image
            last_preview = preview_latent
            down_res, mid_res = aggregator(image, t, encoder_hidden_states=prompt_embeds,
                                           controlnet_cond=preview_latent,
                                           added_cond_kwargs={"text_embeds": add_text_embeds, "time_ids": time_ids},
                                           return_dict=False)
        down_scaled = [s * cond_scale for s in down_res]
        mid_scaled = mid_res * cond_scale
        noise_pred = unet(x_in, t, encoder_hidden_states=prompt_embeds, cross_attention_kwargs=ca_kwargs,
                          down_block_additional_residuals=down_scaled, mid_block_additional_residual=mid_scaled,
                          added_cond_kwargs=added, return_dict=False)[0]
        if do_cfg:
            e_u, e_c = noise_pred.chunk(2)
            noise_pred = e_u + guidance_scale * (e_c - e_u)
            if guidance_rescale > 0.0:  # pipelines/sdxl_instantir.py:1623-1625
                noise_pred = rescale_noise_cfg(noise_pred, e_c, guidance_rescale)
        out = scheduler.step(noise_pred, t, latents, generator=generator, return_dict=True)
        latents = out.prev_sample
        if adastep_restore:  # :1636-1644 (preview_latent is the CFG-concatenated tensor: the cond half is its tail)
            if last_preview is None:
                raise RuntimeError("adastep_restore before any controlled step (the reference raises NameError here)")
            pv = last_preview[latents.shape[0]:] if do_cfg else last_preview
            pred_x0_l2 = (pv.float() - out.pred_original_sample.float()).pow(2).sum(dim=(1, 2, 3))
            previewer_l2 = (pv.float() - previewer_mean.float()).pow(2).sum(dim=(1, 2, 3))
            previewer_mean = pv
            preview_factor = (pred_x0_l2 / previewer_l2).reshape(-1, 1, 1, 1)
        if record is not None:
            record.setdefault("latents", []).append(latents.clone())
            record.setdefault("pred_x0", []).append(out.pred_original_sample.clone())
            record.setdefault("noise_pred", []).append(noise_pred.clone())
            record.setdefault("preview_factor", []).append(preview_factor.reshape(-1).clone())
            record.setdefault("preview", []).append(None if preview_latent is None else preview_latent.clone())
    return latents


def scheduler_add_noise(scheduler, x0, noise, t):
    ac = scheduler.alphas_cumprod.to(dtype=x0.dtype)
    t = int(t)
    return ac[t] ** 0.5 * x0 + (1 - ac[t]) ** 0.5 * noise
