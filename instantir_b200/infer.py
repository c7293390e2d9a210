"""Command-line restoration — the surface of the reference's ``infer.py`` (flags :229-386, batch loop :114-225) on the
sm_100a pipeline.

    python -m instantir_b200.infer --sdxl_path <dir> --instantir_path <dir> --vision_encoder_path <dir> \\
        --test_path <file|dir> --out_path ./output [--cfg 7.0 --preview_start 0.0 --creative_start 1.0 ...]

Same flag names and meaning (``--cfg`` -> guidance_scale, ``--preview_start``, ``--creative_start`` ->
control_guidance_end, infer.py:218-221), same input resizing (``resize_img`` :31-66), same output naming and the same
skip-if-already-written resume (:151-162).  What differs:

* multi-GPU: under ``torchrun`` the batches are dealt round-robin to the ranks (data parallel, no communication);
* images of one batch that resize to different runtime sizes are run as separate sub-batches (the reference's
  image processor silently resizes them all to the first image's size);
* ``--random_init`` (no checkpoints on the box: smoke tests, benchmarks) builds every model with seeded random weights
  and needs no model directory; prompts are then hashed into token ids instead of tokenised.

Checkpoints are read from the layouts the reference reads: a diffusers SDXL directory (``unet/``, ``vae/``,
``text_encoder/``, ``text_encoder_2/``, ``tokenizer/``, ``tokenizer_2/``, ``scheduler/``), ``<instantir_path>/adapter.pt``,
``aggregator.pt``, ``previewer_lora_weights.bin`` and a DINOv2 directory.  File I/O and tokenisation are host glue (PIL,
safetensors, the user's ``transformers`` tokenizers); all arithmetic runs in the kernels.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
from typing import List, Optional, Tuple

import torch

DEFAULT_PROMPT = ("Photorealistic, highly detailed, hyper detailed photo - realistic maximum detail, 32k, "
                  "ultra HD, extreme meticulous detailing, skin pore detailing, hyper sharpness, perfect without deformations, "
                  "taken using a Canon EOS R camera, Cinematic, High Contrast, Color Grading. ")
DEFAULT_NEG_PROMPT = ("blurry, out of focus, unclear, depth of field, over-smooth, sketch, oil painting, cartoon, CG Style, "
                      "3D render, unreal engine, dirty, messy, worst quality, low quality, frames, painting, illustration, drawing, art, "
                      "watermark, signature, jpeg artifacts, deformed, lowres")


def runtime_size(w: int, h: int, max_side=1024, min_side=768, width=None, height=None, base_pixel_number=64) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """the size arithmetic of the reference's ``resize_img`` (infer.py:31-66): ((runtime w, h), (output w, h)).
    Output size = requested width/height (aspect kept when only one is given); runtime size = output size scaled so that
    min side >= min_side, then max side <= max_side, floored to multiples of 64."""
    if width is not None and height is not None:
        out_w, out_h = width, height
    elif width is not None:
        out_w, out_h = width, round(h * width / w)
    elif height is not None:
        out_h, out_w = height, round(w * height / h)
    else:
        out_w, out_h = w, h
    w, h = out_w, out_h
    if min(w, h) < min_side:
        ratio = min_side / min(w, h)
        w, h = round(ratio * w), round(ratio * h)
    if max(w, h) > max_side:
        ratio = max_side / max(w, h)
        w, h = round(ratio * w), round(ratio * h)
    return ((w // base_pixel_number) * base_pixel_number, (h // base_pixel_number) * base_pixel_number), (out_w, out_h)


def resize_img(input_image, max_side=1024, min_side=768, width=None, height=None, pad_to_max_side=False, mode=None, base_pixel_number=64):
    """infer.py:31-66 on a PIL image: (resized image, (out_w, out_h))"""
    import numpy as np
    from PIL import Image

    (rw, rh), out = runtime_size(*input_image.size, max_side, min_side, width, height, base_pixel_number)
    input_image = input_image.resize([rw, rh], Image.BILINEAR if mode is None else mode)
    if pad_to_max_side:
        res = np.ones([max_side, max_side, 3], dtype=np.uint8) * 255
        ox, oy = (max_side - rw) // 2, (max_side - rh) // 2
        res[oy:oy + rh, ox:ox + rw] = np.array(input_image)
        input_image = Image.fromarray(res)
    return input_image, out


def plan_batches(all_inputs: List[str], processed: List[str], batch_size: int, rank: int = 0, world: int = 1) -> List[List[str]]:
    """infer.py:151-169: sorted inputs minus the files already present in the output directory, cut into batches of
    `batch_size` (the last one may be short); rank r of a data-parallel job takes batches r, r + world, ..."""
    todo = [f for f in sorted(all_inputs) if f not in set(processed)]
    batches = [todo[i:i + batch_size] for i in range(0, len(todo), batch_size)]
    return batches[rank::world]


def hashed_token_ids(texts: List[str], vocab: int = 49408, length: int = 77) -> torch.Tensor:
    """--random_init only: a deterministic stand-in for the tokenizer (no vocabulary files on the box): BOS, one id per
    whitespace-separated word from a hash, EOS (the highest id: the pooled position of the legacy eos rule), padding"""
    ids = torch.zeros(len(texts), length, dtype=torch.int64)
    for r, t in enumerate(texts):
        words = t.split()[:length - 2]
        row = [vocab - 2] + [int.from_bytes(hashlib.sha256(w.encode()).digest()[:4], "little") % (vocab - 2) for w in words] + [vocab - 1]
        ids[r, :len(row)] = torch.tensor(row)
    return ids


# ------------------------------------------------------------------------------------------- checkpoints
def load_state_dict_file(path: str):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def _first_existing(directory: str, names: List[str]) -> str:
    for n in names:
        p = os.path.join(directory, n)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(f"none of {names} under {directory}")


def convert_previewer_lora(sd) -> dict:
    """previewer_lora_weights.bin (diffusers LoRA layout, keys prefixed 'unet.') -> '<module>.lora_{A,B}.weight'
    (what pipelines/sdxl_instantir.py:356-370 does with convert_unet_state_dict_to_peft); returns (lora dict, alpha)"""
    out = {}
    for k, v in sd.items():
        if not k.startswith("unet."):
            continue
        k = k[len("unet."):]
        for a, b in ((".lora.down.weight", ".lora_A.weight"), (".lora.up.weight", ".lora_B.weight"),
                     (".lora_linear_layer.down.weight", ".lora_A.weight"), (".lora_linear_layer.up.weight", ".lora_B.weight"),
                     (".lora_A.weight", ".lora_A.weight"), (".lora_B.weight", ".lora_B.weight")):
            if k.endswith(a):
                k = k[: -len(a)] + b
                break
        # processor sub-modules live under '<attn2>.processor.' in this build's key layout (:365-368)
        for name in ("to_k_ip", "to_v_ip", "ln_k_ip", "ln_v_ip"):
            k = k.replace(f"attn2.{name}", f"attn2.processor.{name}")
        out[k] = v
    return out


def revise_adapter_state_dict(sd) -> dict:
    """module/ip_adapter/utils.py:84-99,164-177: adapter.pt = {'image_proj': Resampler sd, 'ip_adapter': ModuleList sd} (legacy
    prefixes 'image_proj_model.' / 'adapter_modules.') -> flat UNet-style keys"""
    if "image_proj" not in sd:
        sd = {"image_proj": {k[len("image_proj_model."):]: v for k, v in sd.items() if k.startswith("image_proj_model.")},
              "ip_adapter": {k[len("adapter_modules."):]: v for k, v in sd.items() if k.startswith("adapter_modules.")}}
    return sd


def build_pipeline(args, device):
    from . import config as pcfg, encoders as enc, weights
    from .aggregator import Aggregator
    from .pipeline import InstantIRPipeline
    from .schedulers import DDPMScheduler, LCMSingleStepScheduler
    from .unet import UNet2DConditionModel
    from .vae import AutoencoderKL, VaeConfig, vae_param_shapes

    cfg = pcfg.sdxl()
    if args.random_init:
        ush = weights.unet_param_shapes(cfg, adapter=True)
        unet = UNet2DConditionModel(cfg, weights.RandomSource(ush, device, seed=0, lora_shapes=weights.lora_param_shapes(cfg, ush)), device, args.precision)
        agg = Aggregator(cfg, weights.RandomSource(weights.aggregator_param_shapes(cfg), device, seed=1), device, args.precision)
        vcfg = VaeConfig()
        vae = AutoencoderKL(vcfg, weights.RandomSource(vae_param_shapes(vcfg), device, seed=2), device, "bf16")
        cl, cg, dc = enc.clip_l(), enc.clip_bigg(), enc.Dinov2Config()
        te = enc.CLIPTextModel(cl, weights.RandomSource(enc.clip_text_param_shapes(cl, False), device, seed=3), device, args.precision)
        te2 = enc.CLIPTextModel(cg, weights.RandomSource(enc.clip_text_param_shapes(cg, True), device, seed=4), device, args.precision, with_projection=True)
        dino = enc.Dinov2Model(dc, weights.RandomSource(enc.dinov2_param_shapes(dc), device, seed=5), device, args.precision)
        pipe = InstantIRPipeline(unet, agg, DDPMScheduler(), vae=vae, text_encoder=te, text_encoder_2=te2, image_encoder=dino)
        return pipe, LCMSingleStepScheduler()
    from transformers import CLIPTokenizer

    sdxl = args.sdxl_path
    usd = load_state_dict_file(_first_existing(os.path.join(sdxl, "unet"), ["diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"]))
    adapter = revise_adapter_state_dict(load_state_dict_file(args.adapter_model_path or os.path.join(args.instantir_path, "adapter.pt")))
    usd.update({"encoder_hid_proj.image_projection_layers.0." + k: v for k, v in adapter["image_proj"].items()})
    lora, lora_scale = None, 1.0
    lpath = args.previewer_lora_path or args.instantir_path
    if lpath is not None:
        lfile = lpath if os.path.isfile(lpath) else os.path.join(lpath, "previewer_lora_weights.bin")
        raw = load_state_dict_file(lfile)
        lora = convert_previewer_lora(raw)
        alphas = [float(v) for k, v in raw.items() if k.endswith(".alpha")]
        rank = next(v.shape[0] for k, v in lora.items() if k.endswith("lora_A.weight"))
        lora_scale = (alphas[0] if alphas else float(rank)) / rank
        print(f"use lora alpha {lora_scale * rank}")
    unet = UNet2DConditionModel(cfg, weights.StateDictSource(usd, device, lora=lora, lora_scale=lora_scale), device, args.precision, adapter=False)
    # processors: adapter.pt's ModuleList order is the order of unet.attn_processors (module/ip_adapter/utils.py:145-152)
    from .attention_processor import init_attn_proc
    from .ip_adapter_utils import load_adapter_to_unet

    names = [n for n, _ in unet.attention_modules()]
    for k, v in adapter["ip_adapter"].items():
        idx, rest = k.split(".", 1)
        usd[f"{names[int(idx)]}.processor.{rest}"] = v
    load_adapter_to_unet(unet)
    agg = Aggregator.from_unet(unet)
    agg.load_state_dict(load_state_dict_file(os.path.join(args.instantir_path, "aggregator.pt")))
    vdir = args.pretrained_vae_model_name_or_path or os.path.join(sdxl, "vae")
    vae = AutoencoderKL(VaeConfig(), weights.StateDictSource(load_state_dict_file(_first_existing(vdir, ["diffusion_pytorch_model.fp16.safetensors", "diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.bin"])), device), device, "bf16")
    te_sd = load_state_dict_file(_first_existing(os.path.join(sdxl, "text_encoder"), ["model.fp16.safetensors", "model.safetensors", "pytorch_model.bin"]))
    te2_sd = load_state_dict_file(_first_existing(os.path.join(sdxl, "text_encoder_2"), ["model.fp16.safetensors", "model.safetensors", "pytorch_model.bin"]))
    dino_sd = load_state_dict_file(_first_existing(args.vision_encoder_path, ["model.safetensors", "pytorch_model.bin"]))
    te = enc.CLIPTextModel(enc.clip_l(), weights.StateDictSource(te_sd, device), device, args.precision)
    te2 = enc.CLIPTextModel(enc.clip_bigg(), weights.StateDictSource(te2_sd, device), device, args.precision, with_projection=True)
    dino = enc.Dinov2Model(enc.Dinov2Config(), weights.StateDictSource(dino_sd, device), device, args.precision)
    sched_cfg = {}
    sp = os.path.join(sdxl, "scheduler", "scheduler_config.json")
    if os.path.exists(sp):
        sched_cfg = {k: v for k, v in json.load(open(sp)).items() if not k.startswith("_")}
    pipe = InstantIRPipeline(unet, agg, DDPMScheduler.from_config(sched_cfg), vae=vae, text_encoder=te, text_encoder_2=te2,
                             tokenizer=CLIPTokenizer.from_pretrained(sdxl, subfolder="tokenizer"),
                             tokenizer_2=CLIPTokenizer.from_pretrained(sdxl, subfolder="tokenizer_2"), image_encoder=dino)
    pipe.prepare_previewers()
    return pipe, LCMSingleStepScheduler.from_config(sched_cfg)


# ------------------------------------------------------------------------------------------------ main
def pil_to_tensor(img) -> torch.Tensor:
    import numpy as np

    return torch.from_numpy(np.asarray(img, dtype=np.float32) / 127.5 - 1.0).permute(2, 0, 1).contiguous()


def main(args, device):
    from PIL import Image

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    pipe, lcm_scheduler = build_pipeline(args, device)
    post_fix = f"_{args.post_fix}" if args.post_fix else ""
    out_dir = os.path.join(args.out_path, post_fix)
    os.makedirs(out_dir, exist_ok=True)
    single = os.path.isfile(args.test_path)
    all_inputs = [os.path.basename(args.test_path)] if single else os.listdir(args.test_path)
    batches = plan_batches(all_inputs, os.listdir(out_dir), args.batch_size, rank, world)
    for f in sorted(set(all_inputs) & set(os.listdir(out_dir))):
        print(f"Skip {f}")
    n_done = 0
    for lq_batch in batches:
        generator = torch.Generator(device=device).manual_seed(args.seed)
        items = []
        for name in lq_batch:
            pil = Image.open(args.test_path if single else os.path.join(args.test_path, name)).convert("RGB")
            pil, out_size = resize_img(pil, width=args.width, height=args.height)
            items.append((name, pil, out_size))
        timesteps = None
        if args.denoising_start < 1000:
            # infer.py:184-190 builds a custom list and then overwrites it with the scheduler's own; mirrored as written
            pipe.scheduler.set_timesteps(args.num_inference_steps)
            timesteps = [int(t) for t in pipe.scheduler.timesteps]
        for size in sorted({it[1].size for it in items}):  # one sub-batch per runtime size
            group = [it for it in items if it[1].size == size]
            lq = torch.stack([pil_to_tensor(it[1]) for it in group])
            prompt = args.prompt if args.prompt else DEFAULT_PROMPT
            prompt = (prompt if isinstance(prompt, list) else [prompt])
            prompt = prompt * len(group) if len(prompt) == 1 else prompt[:len(group)]
            neg = args.neg_prompt if args.neg_prompt else DEFAULT_NEG_PROMPT
            neg = (neg if isinstance(neg, list) else [neg])
            neg = neg * len(group) if len(neg) == 1 else neg[:len(group)]
            if args.random_init:
                prompt, neg = hashed_token_ids(prompt), hashed_token_ids(neg)
            images = pipe(prompt=prompt, image=lq, num_inference_steps=None if timesteps else args.num_inference_steps, generator=generator,
                          timesteps=timesteps, negative_prompt=neg, guidance_scale=args.cfg, previewer_scheduler=lcm_scheduler,
                          preview_start=args.preview_start, control_guidance_end=args.creative_start, output_type="pt").images
            for (name, _, out_size), img in zip(group, images):
                arr = (img.clamp(0, 1).permute(1, 2, 0).cpu().numpy() * 255).round().astype("uint8")
                Image.fromarray(arr).resize([out_size[0], out_size[1]], Image.BILINEAR).save(os.path.join(out_dir, name))
                n_done += 1
    return n_done


def build_parser():
    p = argparse.ArgumentParser(description="InstantIR pipeline (sm_100a)")
    p.add_argument("--sdxl_path", type=str, default=None, help="diffusers SDXL directory")
    p.add_argument("--previewer_lora_path", type=str, default=None, help="Path to the previewer LoRA (default: --instantir_path)")
    p.add_argument("--pretrained_vae_model_name_or_path", type=str, default=None)
    p.add_argument("--instantir_path", type=str, default=None, help="directory with adapter.pt, aggregator.pt, previewer_lora_weights.bin")
    p.add_argument("--vision_encoder_path", type=str, default=None, help="DINOv2-large directory")
    p.add_argument("--adapter_model_path", type=str, default=None)
    p.add_argument("--adapter_tokens", type=int, default=64)
    p.add_argument("--use_clip_encoder", action="store_true")
    p.add_argument("--denoising_start", type=int, default=1000)
    p.add_argument("--num_inference_steps", type=int, default=30)
    p.add_argument("--creative_start", type=float, default=1.0)
    p.add_argument("--preview_start", type=float, default=0.0)
    p.add_argument("--resolution", type=int, default=1024)
    p.add_argument("--batch_size", type=int, default=6)
    p.add_argument("--width", type=int, default=None)
    p.add_argument("--height", type=int, default=None)
    p.add_argument("--cfg", type=float, default=7.0)
    p.add_argument("--post_fix", type=str, default=None)
    p.add_argument("--variant", type=str, default="fp16")
    p.add_argument("--revision", type=str, default=None)
    p.add_argument("--prompt", type=str, default="", nargs="+")
    p.add_argument("--neg_prompt", type=str, default="", nargs="+")
    p.add_argument("--test_path", type=str, default=None, required=True)
    p.add_argument("--out_path", type=str, default="./output")
    p.add_argument("--seed", type=int, default=42)
    # extensions
    p.add_argument("--precision", type=str, default="fp16", choices=["fp16", "bf16", "fp32"])
    p.add_argument("--random_init", action="store_true", help="seeded random weights for every model (no checkpoint directories needed)")
    return p


def cli(argv: Optional[List[str]] = None):
    args = build_parser().parse_args(argv)
    if args.use_clip_encoder:
        raise NotImplementedError("--use_clip_encoder (CLIP-vision image encoder) is outside this build's scope; InstantIR ships with DINOv2")
    if not args.random_init and (args.sdxl_path is None or args.instantir_path is None or args.vision_encoder_path is None):
        raise SystemExit("--sdxl_path, --instantir_path and --vision_encoder_path are required (or --random_init)")
    if not torch.cuda.is_available():
        raise SystemExit("instantir_b200 runs on CUDA only (the reference's CPU fallback, gradio_demo/app.py:46-49, is not reproduced)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.set_grad_enabled(False)
    n = main(args, torch.device(f"cuda:{local}"))
    print(f"restored {n} image(s)")


if __name__ == "__main__":
    cli()
