"""``InstantIRPipeline`` — the reference's denoising loop (pipelines/sdxl_instantir.py:1385-1666) on
the sm_100a kernels, with the reference's call surface for this path: ``guidance_scale`` (--cfg),
``preview_start``, ``control_guidance_end`` (--creative_start, infer.py:218-221),
``previewer_scheduler``, ``num_inference_steps``, ``generator``, ``timesteps``,
``controlnet_conditioning_scale``, ``init_latents_with_lq``, ``save_preview_row``...

Scope: the per-timestep step, plus the "next" rows f1 (VAE) and f2 (CLIP text + DINOv2 encoders,
instantir_b200/encoders.py).  Conditioning comes either as the encoders' OUTPUTS (``prompt_embeds`` /
``pooled_prompt_embeds`` + negatives, ``ip_adapter_image_embeds``) or, when the pipeline was given the encoders, as
``prompt`` / ``negative_prompt`` (strings through the user's tokenizers, or token-id tensors) and ``ip_adapter_image``
(defaults to the LQ image, :1278-1279).  ``image`` is a 4-channel LQ latent (the reference accepts latents there too,
:1370-1382) or, when the pipeline was given ``vae=``, a 3-channel image in [-1, 1] that is encoded here; the result
is latents (``output_type="latent"``) or decoded images ("pt" / "np").

What changes versus the reference's loop (same results, fewer launches):
  * which of the three step shapes runs at step i (previewer+aggregator+UNet / aggregator+UNet /
    UNet only) is known on the host from controlnet_keep/previewing (:1415-1421) — no `.item()` sync;
  * each model forward is captured once as a CUDA graph and replayed; timestep enters through a
    device scalar;
  * cond_scale multiplication, residual injection and torch.cat are one kernel inside the UNet;
  * CFG combine + DDPM update are one kernel; the LCM preview step is one kernel;
  * CFG-parallel (2 ranks: uncond / cond branch, one all-gather of eps per step) and data-parallel
    sharding are provided by instantir_b200.parallel.
"""
from __future__ import annotations

import gc
from types import SimpleNamespace
from typing import List, Optional

import torch

from . import _lib, ops
from .nn import FMap
from .schedulers import _randn


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None):
    """pipelines/sdxl_instantir.py:195-237."""
    if timesteps is not None:
        scheduler.set_timesteps(timesteps=timesteps, device=device)
        return scheduler.timesteps, len(scheduler.timesteps)
    scheduler.set_timesteps(num_inference_steps, device=device)
    return scheduler.timesteps, num_inference_steps


def step_masks(n_steps, preview_start, preview_end, control_guidance_start, control_guidance_end):
    """controlnet_keep / previewing (pipelines/sdxl_instantir.py:1415-1421)."""
    keep, prev = [], []
    for i in range(n_steps):
        keep.append(1.0 - float(i / n_steps < control_guidance_start or (i + 1) / n_steps > control_guidance_end))
        prev.append(1.0 - float(i / n_steps < preview_start or (i + 1) / n_steps > preview_end))
    return keep, prev


class LaunchCounter:
    """kernels launched by this library, including those replayed through CUDA graphs (the C-side
    counter only sees launches made while capturing)."""

    replayed = 0

    @classmethod
    def total(cls):
        return _lib.launch_count() + cls.replayed


class _Graphed:
    """fn() captured once as a CUDA graph over static tensors, replayed afterwards."""

    def __init__(self, fn, enabled: bool):
        self.fn, self.enabled, self.graph, self.out, self.n_launch = fn, enabled, None, None, 0

    def __call__(self):
        if not self.enabled:
            return self.fn()
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside capture (lazy init, context caches)
                self.fn()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            # torch.cuda.graph captures in cudaStreamCaptureModeGlobal, and since torch 2.9 it no longer runs the garbage
            # collector first.  A dead reference cycle that still owns CUDAGraph objects (the step closures of an earlier
            # shape's graph set) collected DURING this capture would run cudaGraphExecDestroy, which that mode prohibits:
            # the capture is invalidated ("operation not permitted when stream is capturing (function reset)", seen on one
            # rank of the 8-GPU run profiles/bench_r02c_n8_with_partitions.json).  So: collect now, and keep the cycle
            # collector off until the capture has ended.
            if not torch.cuda.is_current_stream_capturing():
                gc.collect()
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    self.out = self.fn()
            finally:
                if gc_was_enabled:
                    gc.enable()
            self.n_launch = _lib.launch_count() - before
            self.graph = g
        self.graph.replay()
        LaunchCounter.replayed += self.n_launch
        return self.out


class InstantIRPipeline:
    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]  # pipelines/sdxl_instantir.py:301

    def __init__(self, unet, aggregator, scheduler, vae=None, text_encoder=None, text_encoder_2=None, tokenizer=None,
                 tokenizer_2=None, feature_extractor=None, image_encoder=None, force_zeros_for_empty_prompt=True):
        self.unet, self.aggregator, self.scheduler = unet, aggregator, scheduler
        self.vae, self.text_encoder, self.text_encoder_2, self.image_encoder = vae, text_encoder, text_encoder_2, image_encoder
        self.tokenizer, self.tokenizer_2, self.feature_extractor = tokenizer, tokenizer_2, feature_extractor
        self.config = SimpleNamespace(force_zeros_for_empty_prompt=force_zeros_for_empty_prompt)
        self.device = unet.rt.device
        self._graphs = {}
        self._zero_image_embeds = {}

    # ------------------------------------------------------------------- once-per-image encoders
    def _token_ids(self, text, tokenizer, max_length=None):
        """`text`: str / list of str (needs the HF-style tokenizer the pipeline was given) or an int64 tensor of ids"""
        if torch.is_tensor(text):
            return text.to(torch.int64)
        if tokenizer is None:
            raise ValueError("string prompts need `tokenizer` / `tokenizer_2` (HF CLIPTokenizer objects; tokenisation is host-side "
                             "string processing), or pass token-id tensors / `prompt_embeds`")
        return tokenizer(text, padding="max_length", max_length=max_length or tokenizer.model_max_length, truncation=True,
                         return_tensors="pt").input_ids

    def encode_prompt(self, prompt, prompt_2=None, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True,
                      negative_prompt=None, negative_prompt_2=None, prompt_embeds=None, negative_prompt_embeds=None,
                      pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, lora_scale=None, clip_skip=None):
        """pipelines/sdxl_instantir.py:400-632 on the sm_100a CLIP encoders (instantir_b200.encoders.CLIPTextModel): both
        text encoders' penultimate hidden states concatenated along the feature axis ([B, 77, 768 + 1280]) and the pooled,
        projected output of the second one; zeros for an absent negative prompt when force_zeros_for_empty_prompt."""
        if isinstance(prompt, str):
            prompt = [prompt]
        batch_size = len(prompt) if prompt is not None else prompt_embeds.shape[0]
        tokenizers = [self.tokenizer, self.tokenizer_2] if self.text_encoder is not None else [self.tokenizer_2]
        encoders = [self.text_encoder, self.text_encoder_2] if self.text_encoder is not None else [self.text_encoder_2]
        if prompt_embeds is None or (do_classifier_free_guidance and negative_prompt_embeds is None and negative_prompt is not None):
            if self.text_encoder_2 is None:
                raise ValueError("encoding prompts needs `text_encoder_2` (and usually `text_encoder`): "
                                 "InstantIRPipeline(..., text_encoder=CLIPTextModel(...), text_encoder_2=CLIPTextModel(..., with_projection=True))")

        def run(texts):
            embeds, pooled = [], None
            for text, tok, enc in zip(texts, tokenizers, encoders):
                out = enc(self._token_ids(text, tok), output_hidden_states=True)
                pooled = out[0]  # "we are only ALWAYS interested in the pooled output of the final text encoder" (:526)
                embeds.append(out.hidden_states[-2] if clip_skip is None else out.hidden_states[-(clip_skip + 2)])
            return torch.cat(embeds, dim=-1), pooled

        if prompt_embeds is None:
            prompt_2 = prompt if prompt_2 is None else prompt_2
            prompt_2 = [prompt_2] if isinstance(prompt_2, str) else prompt_2
            prompt_embeds, pooled_prompt_embeds = run([prompt, prompt_2][-len(encoders):])
        zero_out = negative_prompt is None and self.config.force_zeros_for_empty_prompt
        if do_classifier_free_guidance and negative_prompt_embeds is None and zero_out:
            negative_prompt_embeds = torch.zeros_like(prompt_embeds)
            negative_pooled_prompt_embeds = torch.zeros_like(pooled_prompt_embeds)
        elif do_classifier_free_guidance and negative_prompt_embeds is None:
            negative_prompt = negative_prompt if negative_prompt is not None else ""
            negative_prompt_2 = negative_prompt_2 if negative_prompt_2 is not None else negative_prompt
            neg = [batch_size * [n] if isinstance(n, str) else n for n in (negative_prompt, negative_prompt_2)]
            if prompt is not None and not torch.is_tensor(prompt) and not torch.is_tensor(neg[0]) and type(prompt) is not type(neg[0]):
                raise TypeError(f"`negative_prompt` should be the same type to `prompt`, but got {type(neg[0])} != {type(prompt)}.")
            if batch_size != len(neg[0]):
                raise ValueError(f"`negative_prompt` has batch size {len(neg[0])}, but `prompt` has batch size {batch_size}. Please make "
                                 "sure that passed `negative_prompt` matches the batch size of `prompt`.")
            negative_prompt_embeds, negative_pooled_prompt_embeds = run(neg[-len(encoders):])

        def rep(t, pooled=False):
            if t is None or num_images_per_prompt == 1:
                return t
            return t.repeat(1, num_images_per_prompt).view(t.shape[0] * num_images_per_prompt, -1) if pooled else \
                t.repeat(1, num_images_per_prompt, 1).view(t.shape[0] * num_images_per_prompt, t.shape[1], -1)

        return rep(prompt_embeds), rep(negative_prompt_embeds), rep(pooled_prompt_embeds, True), rep(negative_pooled_prompt_embeds, True)

    def encode_image(self, image, device=None, num_images_per_prompt=1, output_hidden_states=None):
        """pipelines/sdxl_instantir.py:635-669, DINO branch: last_hidden_state of the image and of a zero image.  `image`:
        pixel_values [B,3,224,224] (already preprocessed), or anything the pipeline's `feature_extractor` accepts.  The
        zero-image embedding depends on the weights only: it is computed once per batch size and cached."""
        if self.image_encoder is None:
            raise ValueError("encoding the IP-adapter image needs `image_encoder` (instantir_b200.encoders.Dinov2Model)")
        if output_hidden_states:
            raise NotImplementedError("the CLIP-vision hidden-state branch (use_clip_encoder) is outside this build's scope")
        if not torch.is_tensor(image):
            if self.feature_extractor is None:
                raise ValueError("non-tensor images need `feature_extractor` (the AutoImageProcessor of dinov2-large)")
            image = self.feature_extractor(image, return_tensors="pt").pixel_values
        image_embeds = self.image_encoder(image).last_hidden_state.repeat_interleave(num_images_per_prompt, dim=0)
        key = tuple(image.shape)
        if key not in self._zero_image_embeds:
            self._zero_image_embeds[key] = self.image_encoder(torch.zeros_like(image)).last_hidden_state
        return image_embeds, self._zero_image_embeds[key].repeat_interleave(num_images_per_prompt, dim=0)

    def prepare_ip_adapter_image_embeds(self, ip_adapter_image, ip_adapter_image_embeds, device=None, num_images_per_prompt=1,
                                        do_classifier_free_guidance=True):
        """pipelines/sdxl_instantir.py:672-729 for the single IP-adapter of InstantIR: a list holding one tensor
        [2, B, S, D] = (zero-image, image) DINOv2 tokens ([1, B, S, D] without CFG)."""
        if ip_adapter_image_embeds is not None:
            return ip_adapter_image_embeds if isinstance(ip_adapter_image_embeds, list) else [ip_adapter_image_embeds]
        if isinstance(ip_adapter_image, list):
            if len(ip_adapter_image) != 1:
                raise ValueError(f"`ip_adapter_image` must have same length as the number of IP Adapters. Got {len(ip_adapter_image)} images and 1 IP Adapters.")
            ip_adapter_image = ip_adapter_image[0]
        pos, neg = self.encode_image(ip_adapter_image, device, 1)
        pos, neg = pos.unsqueeze(0), neg.unsqueeze(0)
        return [torch.cat([neg, pos]) if do_classifier_free_guidance else pos]

    def prepare_previewers(self, previewer_lora_path=None, use_lcm=False):
        """Reference: loads previewer_lora_weights.bin into a peft adapter then disables it (:350-397).
        Here the LoRA is part of the UNet's weight source (merged into a second weight set at pack
        time); this only validates that it is present and leaves the adapter disabled."""
        if not any(a.to_out[0].w.lora is not None for _, a in self.unet.attention_modules()):
            raise ValueError("the UNet was built from a weight source without previewer LoRA tensors")
        self.unet.disable_adapters()
        return None

    # ------------------------------------------------------------------------------------ checks
    def check_inputs(self, image, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds,
                     negative_pooled_prompt_embeds, ip_adapter_image_embeds, guidance_scale,
                     control_guidance_start, control_guidance_end, previewer_scheduler, preview_start):
        if prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        if pooled_prompt_embeds is None:
            raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed.")
        if guidance_scale > 1.0 and (negative_prompt_embeds is None or negative_pooled_prompt_embeds is None):
            raise ValueError("classifier-free guidance needs `negative_prompt_embeds` and `negative_pooled_prompt_embeds`.")
        if negative_prompt_embeds is not None and prompt_embeds.shape != negative_prompt_embeds.shape:
            raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, "
                             f"but got {prompt_embeds.shape} != {negative_prompt_embeds.shape}.")
        if not torch.is_tensor(image) or image.ndim != 4:
            raise TypeError("`image` must be a tensor: a 4-channel LQ latent [B,4,h,w] or, with a VAE, an image [B,3,H,W] in [-1,1]")
        if image.shape[1] != self.unet.cfg.in_channels and (self.vae is None or getattr(self.vae, "encoder", None) is None
                                                             or image.shape[1] != self.vae.config.in_channels):
            raise TypeError("`image` must be a 4-channel latent tensor [B,4,h,w] (pass vae=AutoencoderKL(...) with encoder "
                            "weights to give a 3-channel image instead)")
        if ip_adapter_image_embeds is None:
            raise ValueError("Provide `ip_adapter_image` / a 3-channel `image` (with image_encoder=Dinov2Model) or `ip_adapter_image_embeds` "
                             "(DINOv2 tokens [2,B,S,D] or a list holding that tensor).")
        if control_guidance_start >= control_guidance_end:
            raise ValueError(f"control guidance start: {control_guidance_start} cannot be larger or equal to control guidance end: {control_guidance_end}.")
        if control_guidance_start < 0.0:
            raise ValueError(f"control guidance start: {control_guidance_start} can't be smaller than 0.")
        if control_guidance_end > 1.0:
            raise ValueError(f"control guidance end: {control_guidance_end} can't be larger than 1.0.")
        if preview_start < 1.0 and previewer_scheduler is None:
            raise ValueError("previewing steps need `previewer_scheduler` (LCMSingleStepScheduler)")

    # -------------------------------------------------------------------------------------- call
    @torch.no_grad()
    def __call__(self, prompt=None, prompt_2=None, image=None, height=None, width=None, num_inference_steps=30,
                 timesteps: Optional[List[int]] = None, denoising_end=None, guidance_scale=7.0, negative_prompt=None,
                 negative_prompt_2=None, num_images_per_prompt=1, eta=0.0, generator=None, latents=None,
                 prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                 negative_pooled_prompt_embeds=None, ip_adapter_image=None, ip_adapter_image_embeds=None,
                 output_type="latent", return_dict=True, cross_attention_kwargs=None, guidance_rescale=0.0,
                 original_size=None, crops_coords_top_left=(0, 0), target_size=None, save_preview_row=False,
                 init_latents_with_lq=True, multistep_restore=False, adastep_restore=False, previewer_scheduler=None,
                 preview_start=0.0, preview_end=1.0, control_guidance_start=0.0, control_guidance_end=1.0,
                 controlnet_conditioning_scale=1.0, reference_latents=None, use_cuda_graph=True, cfg_parallel=None,
                 dp_shard=None, record=None, overlap_streams=True, agg_ahead=False, **kwargs):
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make sure to only forward one of the two.")
        do_cfg_ = guidance_scale > 1.0
        if prompt is not None:
            # 3.1 Encode input prompt (:1327-1346)
            prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds = self.encode_prompt(
                prompt, prompt_2, None, num_images_per_prompt, do_cfg_, negative_prompt, negative_prompt_2,
                negative_prompt_embeds=negative_prompt_embeds, negative_pooled_prompt_embeds=negative_pooled_prompt_embeds,
                clip_skip=kwargs.get("clip_skip"))
        if ip_adapter_image_embeds is None and (ip_adapter_image is not None or (torch.is_tensor(image) and image.ndim == 4 and image.shape[1] == 3)):
            # 3.2 Encode ip_adapter_image (:1349-1356); the IP image defaults to the LQ input itself (:1278-1279)
            src_img = ip_adapter_image
            if src_img is None:
                from .encoders import dinov2_preprocess

                src_img = dinov2_preprocess((image.float() + 1.0) / 2.0)
            ip_adapter_image_embeds = self.prepare_ip_adapter_image_embeds(src_img, None, None, num_images_per_prompt, True)
        if multistep_restore:
            raise NotImplementedError("multistep_restore calls scheduler.step with kwargs stock DDPM does not have: not runnable in the "
                                      "reference as shipped (SURVEY App. E, §8 f4)")
        if reference_latents is not None and agg_ahead:
            raise NotImplementedError("reference_latents with agg_ahead: the one-step-ahead Aggregator is fed the LQ latent")
        if adastep_restore and (cfg_parallel is not None or agg_ahead):
            raise NotImplementedError("adastep_restore with cfg_parallel / agg_ahead: the adaptive factor lives on the cond rank and "
                                      "makes the Aggregator's schedule data dependent")
        if output_type != "latent":
            if self.vae is None:
                raise ValueError("output_type != 'latent' needs a VAE: InstantIRPipeline(unet, aggregator, scheduler, vae=AutoencoderKL(...))")
            if output_type not in ("pt", "np"):
                raise NotImplementedError(f"output_type={output_type!r}: 'latent', 'pt' and 'np' are built (SURVEY §8 f1)")
        self.check_inputs(image, prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds,
                          negative_pooled_prompt_embeds, ip_adapter_image_embeds, guidance_scale,
                          control_guidance_start, control_guidance_end, previewer_scheduler, preview_start)
        dev = self.device
        unet, agg, sched = self.unet, self.aggregator, self.scheduler
        # dp_shard=(total_images, slice): this rank restores images[slice] of a data-parallel job; noise is
        # drawn for the FULL batch and sliced so the sharded run is bit-comparable with the unsharded one
        if dp_shard is not None:
            if generator is None:
                raise ValueError("dp_shard needs an explicitly seeded `generator` (identical on every rank): each rank draws the "
                                 "full-batch noise and keeps its slice, which is only consistent when the streams agree")
            total_b, sl = dp_shard

            def draw(shape):
                return _randn((total_b,) + tuple(shape[1:]), generator, dev)[sl].contiguous()
        else:
            def draw(shape):
                return _randn(tuple(shape), generator, dev)
        do_cfg = guidance_scale > 1.0
        f32 = dict(device=dev, dtype=torch.float32)
        image = image.to(**f32).contiguous()
        if image.shape[1] != unet.cfg.in_channels:
            # pipelines/sdxl_instantir.py:1370-1376: image -> vae.encode(...).latent_dist.sample() * scaling_factor
            # (the reference draws this sample from torch's global RNG; here from `generator`, before any other draw)
            image = self.vae.encode(image).latent_dist.sample(generator, scale=self.vae.config.scaling_factor)
        B, _, h, w = image.shape
        # :1369,1381-1382: height / width are taken from the prepared image (the `height` / `width` arguments only
        # steer the reference's PIL preprocessing, which tensors bypass)
        H_px, W_px = h * 8, w * 8
        ts, num_inference_steps = retrieve_timesteps(sched, num_inference_steps, dev, timesteps)
        n = len(ts)
        n_run = n
        if denoising_end is not None and isinstance(denoising_end, float) and 0 < denoising_end < 1:
            # 8.1 (:1469-1484): only the timesteps at or above the cut-off run; controlnet_keep / previewing (:1415-1421) and
            # the scheduler keep the FULL schedule
            n_train = sched.config.num_train_timesteps
            cutoff = int(round(n_train - denoising_end * n_train))
            n_run = len([t for t in torch.as_tensor(ts).tolist() if t >= cutoff])
        # 6. latents (:1388-1401): the LQ latent noised to t0 with the user's generator (init_latents :931-939; a passed
        # `latents` is ignored on this branch, as in the reference), else prepare_latents (:942-963)
        if init_latents_with_lq:
            latents = sched.add_noise(image, draw(image.shape), ts[:1])
        elif latents is not None:
            latents = latents.to(**f32).contiguous() * sched.init_noise_sigma
        else:
            latents = draw(image.shape) * sched.init_noise_sigma
        # DDPM variance noise of every step, drawn now in step order (the generator sees exactly the calls the
        # reference's loop makes, :1629): no RNG launch is left inside the loop, and a CFG pair needs one broadcast
        step_noise = [draw(latents.shape) if int(t) > 0 else None for t in ts]
        if cfg_parallel is not None:
            # both ranks of a pair MUST step with identical x_t and z: the pair leader's draws win,
            # whatever generator state the other rank came with
            image = cfg_parallel.broadcast_from_leader(image)
            latents = cfg_parallel.broadcast_from_leader(latents)
            live = [z for z in step_noise if z is not None]
            if live:
                stack = cfg_parallel.broadcast_from_leader(torch.stack(live))
                it = iter(stack.unbind(0))
                step_noise = [next(it) if z is not None else None for z in step_noise]
        keep, previewing = step_masks(n, preview_start, preview_end, control_guidance_start, control_guidance_end)
        scales = controlnet_conditioning_scale if isinstance(controlnet_conditioning_scale, list) else [controlnet_conditioning_scale] * n
        if len(scales) != n:
            raise ValueError(f"{len(scales)} controlnet scales do not match number of sampling steps {n}")
        # 7.2 added time ids (:1428-1455)
        original_size = original_size or (H_px, W_px)
        target_size = target_size or (H_px, W_px)
        time_ids = torch.tensor([list(original_size) + list(crops_coords_top_left) + list(target_size)], **f32).repeat(B, 1)
        if isinstance(ip_adapter_image_embeds, list):
            ip_adapter_image_embeds = ip_adapter_image_embeds[0]
        ip = ip_adapter_image_embeds.to(**f32)
        if ip.ndim != 4 or ip.shape[0] != 2:
            raise ValueError("ip_adapter_image_embeds must be [2, B, S, D] (negative, positive) DINOv2 tokens")
        # branch layout: CFG batch = [uncond; cond] (:1457-1460); a CFG-parallel rank keeps one branch
        branches = [0, 1] if do_cfg else [1]
        if cfg_parallel is not None:
            if not do_cfg:
                raise ValueError("cfg_parallel needs guidance_scale > 1")
            branches = [cfg_parallel.branch]
        pe = {0: negative_prompt_embeds, 1: prompt_embeds}
        pp = {0: negative_pooled_prompt_embeds, 1: pooled_prompt_embeds}
        nb = len(branches) * B
        new = dict(
            prompt_all=torch.cat([pe[b].to(**f32) for b in branches], 0),
            pooled_all=torch.cat([pp[b].to(**f32) for b in branches], 0),
            time_ids_all=time_ids.repeat(len(branches), 1),
            image_all=torch.cat([image] * len(branches), 0),
            ip_all=torch.cat([ip[b] for b in branches], 0).unsqueeze(1))
        ref_given = reference_latents is not None
        if ref_given:
            # :1579-1580: on controlled steps that do not preview, the Aggregator is fed this latent instead of the LQ one
            ref = reference_latents.to(**f32).contiguous()
            if tuple(ref.shape) != tuple(image.shape):
                raise ValueError(f"reference_latents must have the LQ latent's shape {tuple(image.shape)}, got {tuple(ref.shape)}")
            new["ref_all"] = torch.cat([ref] * len(branches), 0)
        # Static tensors + captured graphs are kept across calls of the same shape: new conditioning is
        # copied INTO the static buffers (version bump -> step-invariant caches recompute in place), so the
        # graphs captured for the first image are replayed for every later one.
        key = (tuple(branches), B, h, w, bool(use_cuda_graph), bool(overlap_streams), bool(agg_ahead), agg.weights_version,
               tuple((k, tuple(v.shape)) for k, v in sorted(new.items())))
        S = self._graphs.get(key)
        if S is None:
            S = SimpleNamespace(**{k: v.contiguous().clone() for k, v in new.items()})
            S.x_in = torch.empty(nb, 4, h, w, **f32)
            S.t_dev = torch.empty(1, **f32)
            S.cond_scale = torch.empty(nb, **f32)
            S.preview_latent = torch.empty(nb, 4, h, w, **f32)
            S.previewer_mean = torch.zeros(B, 4, h, w, **f32)  # adastep_restore state (:1488-1492)
            S.preview_factor = torch.ones(B, **f32)
            S.st = SimpleNamespace(down=None, mid=None)
            self._graphs = {key: S}  # one shape at a time: drop graphs of other shapes
            prompt_all, image_all = S.prompt_all, S.image_all
            added = {"text_embeds": S.pooled_all, "time_ids": S.time_ids_all, "image_embeds": [S.ip_all]}
            agg_added = {"text_embeds": S.pooled_all, "time_ids": S.time_ids_all}
            x_in, t_dev, cond_scale, preview_latent, st = S.x_in, S.t_dev, S.cond_scale, S.preview_latent, S.st
            S.added = added

            def f_preview():
                unet.enable_adapters()
                try:
                    return unet(x_in, t_dev, encoder_hidden_states=prompt_all, added_cond_kwargs=added, return_dict=False)[0]
                finally:
                    unet.disable_adapters()

            def f_agg(cond):
                return lambda: agg(image_all, t_dev, encoder_hidden_states=prompt_all, controlnet_cond=cond,
                                   added_cond_kwargs=agg_added, return_dict=False)

            def f_unet(with_res):
                def run():
                    if with_res:
                        return unet(x_in, t_dev, encoder_hidden_states=prompt_all, added_cond_kwargs=added,
                                    down_block_additional_residuals=st.down, mid_block_additional_residual=st.mid,
                                    additional_residual_scale=cond_scale, return_dict=False)[0]
                    return unet(x_in, t_dev, encoder_hidden_states=prompt_all, added_cond_kwargs=added, return_dict=False)[0]
                return run

            S.side = torch.cuda.Stream(device=dev)

            def f_step(cond):
                """Aggregator + UNet as ONE graph with a fork: the UNet's embeddings, down blocks and mid block do
                not read the Aggregator's residuals (they enter the up path), so that half runs on a second
                stream while the Aggregator runs on the first; both are chains of small, under-filled kernels
                and overlap well (tools/bench_streams.py)."""
                def run():
                    cur = torch.cuda.current_stream()
                    S.side.wait_stream(cur)
                    with torch.cuda.stream(S.side):
                        state = unet.forward_down_mid(x_in, t_dev, prompt_all, added_cond_kwargs=added)
                    down, mid = agg(image_all, t_dev, encoder_hidden_states=prompt_all, controlnet_cond=cond,
                                    added_cond_kwargs=agg_added, return_dict=False, head_stream=S.side)
                    cur.wait_stream(S.side)
                    eps = unet.forward_up(state, down, mid, cond_scale)[0]
                    return eps, down, mid
                return run

            # ---- aggregator one step AHEAD.  On a step that does not preview, the Aggregator's inputs are the LQ
            # latent (as sample AND as controlnet_cond, pipelines/sdxl_instantir.py:1578-1593), the prompt and the
            # timestep: nothing that depends on the current latents.  So the Aggregator of step i+1 runs on the side
            # stream beside the WHOLE UNet of step i (not only beside its down path), into the other of two residual
            # sets; the first such step of a run pays one stand-alone Aggregator.  Opt-in (`agg_ahead=True`): measured
            # on B200 it is NOT faster than the in-order fork (37.4 vs 36.0 ms/step) — two tcgen05 GEMM CTAs cannot
            # share an SM (each takes all 512 TMEM columns and > 113 KB of shared memory), so kernels of the two
            # streams mostly take turns whatever the schedule.
            S.t_next = torch.empty(1, **f32)
            S.res_buf = None

            def res_buffers():
                if S.res_buf is None:  # shapes from one eager forward (also the kernels' lazy-init warm-up)
                    down, mid = agg(image_all, t_dev, encoder_hidden_states=prompt_all, controlnet_cond=image_all,
                                    added_cond_kwargs=agg_added, return_dict=False)

                    def like(r):
                        nn_, C_, H_, W_ = r.shape
                        return torch.empty(nn_ * H_ * W_, C_, device=r.device, dtype=r.dtype)

                    S.res_buf = [([like(r) for r in down], like(mid)) for _ in range(2)]
                    S.res_view = [([FMap(b, r.shape[0], r.shape[2], r.shape[3], r.shape[1]).nchw() for b, r in zip(bd, down)],
                                   FMap(bm, mid.shape[0], mid.shape[2], mid.shape[3], mid.shape[1]).nchw()) for bd, bm in S.res_buf]
                return S.res_buf

            def f_agg_into(p):
                def run():
                    agg(image_all, t_dev, encoder_hidden_states=prompt_all, controlnet_cond=image_all,
                        added_cond_kwargs=agg_added, return_dict=False, out_buffers=res_buffers()[p])
                return run

            def f_unet_buf(p):
                def run():
                    res_buffers()
                    return unet(x_in, t_dev, encoder_hidden_states=prompt_all, added_cond_kwargs=added,
                                down_block_additional_residuals=S.res_view[p][0], mid_block_additional_residual=S.res_view[p][1],
                                additional_residual_scale=cond_scale, return_dict=False)[0]
                return run

            def f_pipe(p):
                def run():
                    bufs = res_buffers()
                    cur = torch.cuda.current_stream()
                    S.side.wait_stream(cur)
                    with torch.cuda.stream(S.side):
                        agg(image_all, S.t_next, encoder_hidden_states=prompt_all, controlnet_cond=image_all,
                            added_cond_kwargs=agg_added, return_dict=False, out_buffers=bufs[1 - p])
                    eps = unet(x_in, t_dev, encoder_hidden_states=prompt_all, added_cond_kwargs=added,
                               down_block_additional_residuals=S.res_view[p][0], mid_block_additional_residual=S.res_view[p][1],
                               additional_residual_scale=cond_scale, return_dict=False)[0]
                    cur.wait_stream(S.side)
                    return eps
                return run

            S.g_agg_into = [_Graphed(f_agg_into(p), use_cuda_graph) for p in range(2)]
            S.g_pipe = [_Graphed(f_pipe(p), use_cuda_graph) for p in range(2)]
            S.g_step = {"prev": _Graphed(f_step(preview_latent), use_cuda_graph), "lq": _Graphed(f_step(image_all), use_cuda_graph)}
            S.g_preview = _Graphed(f_preview, use_cuda_graph)
            S.g_agg_prev = _Graphed(f_agg(preview_latent), use_cuda_graph)
            S.g_agg_lq = _Graphed(f_agg(image_all), use_cuda_graph)
            # one graph per residual source: the captured UNet reads the static outputs of that aggregator graph
            S.g_unet_res = {"prev": _Graphed(f_unet(True), use_cuda_graph), "lq": _Graphed(f_unet(True), use_cuda_graph),
                            "buf0": _Graphed(f_unet_buf(0), use_cuda_graph), "buf1": _Graphed(f_unet_buf(1), use_cuda_graph)}
            S.g_unet_plain = _Graphed(f_unet(False), use_cuda_graph)
            S.g_agg_ref = None
            if ref_given:
                S.g_step["ref"] = _Graphed(f_step(S.ref_all), use_cuda_graph)
                S.g_agg_ref = _Graphed(f_agg(S.ref_all), use_cuda_graph)
                S.g_unet_res["ref"] = _Graphed(f_unet(True), use_cuda_graph)
        else:
            for k, v in new.items():
                getattr(S, k).copy_(v)
            S.st.down = S.st.mid = None
        x_in, t_dev, cond_scale, preview_latent, st = S.x_in, S.t_dev, S.cond_scale, S.preview_latent, S.st
        g_preview, g_agg_prev, g_agg_lq, g_unet_res, g_unet_plain = S.g_preview, S.g_agg_prev, S.g_agg_lq, S.g_unet_res, S.g_unet_plain
        g_step = S.g_step if overlap_streams else None
        unet.refresh_context(S.prompt_all, S.added, None)
        if adastep_restore:
            S.previewer_mean.zero_()
            S.preview_factor.fill_(1.0)
            S.cond_scale.fill_(min(1.0, float(scales[0])) * keep[0])
        loop = SimpleNamespace(latents=latents, res_src=None, preview_row=[], n_steps=n_run, timesteps=ts, ahead=None, last_previewed=False)
        ahead_ok = g_step is not None and agg_ahead and not ref_given

        def lq_step(j):
            """step j runs the Aggregator on the LQ latent alone (no preview): its output depends only on t_j"""
            return j < n and min(1.0, float(scales[j])) * keep[j] > 0.1 and not previewing[j] > 0

        def step(i):
            """one denoising step of the schedule (pipelines/sdxl_instantir.py:1497-1666)."""
            t_int = int(ts[i])
            lat = loop.latents
            cs = min(1.0, float(scales[i])) * keep[i]  # preview_factor == 1 without adastep_restore
            if adastep_restore:
                # cond_scale = clamp(preview_factor, 0, scale_i) * keep_i per image (:1538-1540) was written on the device by the
                # previous step's iir_adastep_update; the gate below (:1542) is data dependent, so this option pays the
                # reference's host sync
                ops.step_prologue(lat, x_in, len(branches), t=float(t_int), t_dev=t_dev, cond_scale=0.0, cond_scale_dev=None)
                cs_host = cond_scale.cpu()
                cs = float(cs_host.max()) if bool((cs_host > 0.1).any()) else min(0.1, float(cs_host.max()))
            else:
                # torch.cat([latents]*2) (:1503; scale_model_input = id), t and cond_scale (:1538-1540) in ONE launch
                ops.step_prologue(lat, x_in, len(branches), t=float(t_int), t_dev=t_dev, cond_scale=cs, cond_scale_dev=cond_scale)
            previewed = False
            noise_pred = None
            if cs > 0.1:  # the `(cond_scale>0.1).sum().item() > 0` gate (:1542), decided on the host
                if previewing[i] > 0:
                    preview_noise = g_preview()
                    previewer_scheduler.step(preview_noise, t_int, x_in, return_dict=False, out=preview_latent)
                    previewed = True
                    loop.last_previewed = True
                    # the reference keeps the cond chunk (:1564-1567); in a CFG pair only the cond rank holds it
                    if save_preview_row and (cfg_parallel is None or cfg_parallel.branch == 1):
                        loop.preview_row.append(preview_latent[-B:].clone())
                    loop.res_src = "prev"
                else:
                    loop.res_src = "ref" if ref_given else "lq"
                    loop.last_previewed = False
                if not previewed and ahead_ok:
                    # Aggregator(t_i) was computed beside the UNet of step i-1 (or is computed now, once per run);
                    # Aggregator(t_{i+1}) runs beside this step's UNet
                    if loop.ahead is not None and loop.ahead[0] == i:
                        p = loop.ahead[1]
                    else:
                        p = 0
                        S.g_agg_into[p]()
                    if lq_step(i + 1):
                        S.t_next.fill_(float(int(ts[i + 1])))
                        noise_pred = S.g_pipe[p]()
                        loop.ahead = (i + 1, 1 - p)
                    else:
                        noise_pred = g_unet_res[f"buf{p}"]()
                        loop.ahead = None
                    st.down, st.mid = S.res_view[p]
                    loop.res_src = f"buf{p}"
                elif g_step is not None:
                    loop.ahead = None
                    noise_pred, st.down, st.mid = g_step[loop.res_src]()   # aggregator || UNet down+mid, then UNet up
                else:
                    st.down, st.mid = {"prev": g_agg_prev, "lq": g_agg_lq, "ref": S.g_agg_ref}[loop.res_src]()
            if noise_pred is not None:
                pass
            elif st.down is None:
                if cs > 0:
                    raise RuntimeError("control is active but no aggregator features exist")
                noise_pred = g_unet_plain()  # reference would raise NameError here; UNet-only is the intent
            elif cs == 0.0 and not adastep_restore:
                noise_pred = g_unet_plain()  # stale residuals x 0 (:1602-1603) == no residuals
            else:
                noise_pred = g_unet_res[loop.res_src]()
            if cfg_parallel is not None:
                noise_pred = cfg_parallel.gather_branches(noise_pred)  # [2B,4,h,w] = [uncond; cond]
            g_step_arg = guidance_scale if do_cfg else None
            if do_cfg and guidance_rescale > 0.0:
                # rescale_noise_cfg (:181-192, :1623-1625): needs the per-sample std of the guided and of the text
                # prediction, so the combine leaves the fused CFG+DDPM kernel for a one-CTA-per-sample kernel
                nb_ = noise_pred.shape[0] // 2
                eps32 = noise_pred.float().contiguous()
                noise_pred = ops.cfg_rescale(eps32[:nb_], eps32[nb_:], torch.empty_like(eps32[:nb_]), guidance=guidance_scale,
                                             rescale=guidance_rescale)
                g_step_arg = None
            out = sched.step(noise_pred, t_int, lat, generator=generator, return_dict=True,
                             guidance=g_step_arg,
                             noise=step_noise[i])
            loop.latents = out.prev_sample
            if adastep_restore:
                # :1636-1644 + the next step's clamp (:1538): one launch; `preview` = the cond half of the LAST preview latent
                # the Aggregator was fed (the LQ latent on non-previewing steps)
                if loop.res_src is None:
                    raise RuntimeError("adastep_restore before any controlled step (the reference raises NameError here)")
                pv = preview_latent[-B:] if loop.last_previewed else (S.ref_all if ref_given else S.image_all)[-B:]
                nxt = i + 1 if i + 1 < n else i
                ops.adastep_update(pv, out.pred_original_sample, S.previewer_mean, S.preview_factor, cond_scale, n_rep=len(branches),
                                   next_scale=float(scales[nxt]), next_keep=keep[nxt])
            if record is not None:
                record.setdefault("preview_factor", []).append(S.preview_factor.clone())
                record.setdefault("latents", []).append(loop.latents.clone())
                record.setdefault("pred_x0", []).append(out.pred_original_sample.clone())
                record.setdefault("preview", []).append(preview_latent.clone() if previewed else None)
            return loop.latents

        loop.step = step
        if kwargs.get("prepare_only"):
            return loop
        callback_on_step_end = kwargs.get("callback_on_step_end")
        callback, callback_steps = kwargs.get("callback"), kwargs.get("callback_steps") or 1
        names = kwargs.get("callback_on_step_end_tensor_inputs", ["latents"])
        bad = [k for k in names if k not in self._callback_tensor_inputs]
        if bad:  # check_inputs, :774-779
            raise ValueError(f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found {bad}")
        nbr = len(branches)
        for i in range(n_run):
            step(i)
            if callback_on_step_end is not None:
                # :1650-1658.  Inside the reference's loop `prompt_embeds` is the CFG-concatenated tensor the UNet reads;
                # whatever the callback returns replaces it for the following steps (here: copied into the static
                # conditioning buffer, step-invariant K/V recomputed in place)
                avail = {"latents": loop.latents, "prompt_embeds": S.prompt_all, "negative_prompt_embeds": negative_prompt_embeds}
                outputs = callback_on_step_end(self, i, ts[i], {k: avail[k] for k in names}) or {}
                loop.latents = outputs.pop("latents", loop.latents).to(**f32).contiguous()
                new_pe = outputs.pop("prompt_embeds", None)
                outputs.pop("negative_prompt_embeds", None)  # rebinds a name the reference's loop never reads again
                if new_pe is not None and new_pe is not S.prompt_all:
                    if tuple(new_pe.shape) != tuple(S.prompt_all.shape):
                        raise ValueError(f"callback returned prompt_embeds of shape {tuple(new_pe.shape)}, expected "
                                         f"{tuple(S.prompt_all.shape)} ({nbr} CFG branch(es) x batch)")
                    S.prompt_all.copy_(new_pe.to(**f32))
                    unet.refresh_context(S.prompt_all, S.added, None)
            if callback is not None and i % callback_steps == 0:  # deprecated form (:1660-1664)
                callback(i // getattr(sched, "order", 1), ts[i], loop.latents)
        latents, preview_row = loop.latents, loop.preview_row
        if output_type != "latent":
            # pipelines/sdxl_instantir.py:1670-1725: (latents * std / scaling_factor + mean | latents / scaling_factor) ->
            # vae.decode -> postprocess, for the result and, with save_preview_row, for every stored preview latent
            from .vae import postprocess

            vcfg = self.vae.config
            mean, std = getattr(vcfg, "latents_mean", None), getattr(vcfg, "latents_std", None)

            def to_image(z):
                if mean is not None and std is not None:
                    z = z * torch.tensor(std, **f32).view(1, -1, 1, 1) / vcfg.scaling_factor + torch.tensor(mean, **f32).view(1, -1, 1, 1)
                else:
                    z = z / vcfg.scaling_factor
                return postprocess(self.vae.decode(z, return_dict=False)[0], output_type)

            latents = to_image(latents)
            if save_preview_row:
                preview_row = [to_image(pz) for pz in preview_row]
        if not return_dict:
            return (latents, preview_row) if save_preview_row else (latents,)
        return SimpleNamespace(images=latents, preview_rows=preview_row if save_preview_row else None)
