"""the command-line surface of the reference's infer.py.  CPU part: size arithmetic against the reference's
own ``resize_img`` (executed verbatim by tests/golden/make_golden_misc.py -> infer_cli.json), flag names, batching / resume
plan, checkpoint key conversion.  GPU part: the CLI end to end on random-init weights."""
import json
import os
import subprocess
import sys

import pytest
import torch

from instantir_b200 import infer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_resize_img_matches_reference_run():
    from PIL import Image

    for c in json.load(open(os.path.join(G, "infer_cli.json")))["resize_img"]:
        runtime, out = infer.runtime_size(c["w"], c["h"], width=c["width"], height=c["height"])
        assert list(runtime) == c["runtime"] and list(out) == c["out"], c
        img, out2 = infer.resize_img(Image.new("RGB", (c["w"], c["h"])), width=c["width"], height=c["height"])
        assert list(img.size) == c["runtime"] and list(out2) == c["out"]
        assert img.size[0] % 64 == 0 and img.size[1] % 64 == 0


def test_cli_has_every_reference_flag_with_the_same_defaults():
    ref_flags = json.load(open(os.path.join(G, "infer_cli.json")))["flags"]
    parser = infer.build_parser()
    ours = {o for a in parser._actions for o in a.option_strings}
    assert set(ref_flags) <= ours, sorted(set(ref_flags) - ours)
    d = vars(parser.parse_args(["--test_path", "x"]))
    assert (d["num_inference_steps"], d["cfg"], d["batch_size"], d["preview_start"], d["creative_start"], d["seed"], d["denoising_start"]) == \
        (30, 7.0, 6, 0.0, 1.0, 42, 1000)   # infer.py:286-331,385


def test_batch_plan_skips_processed_files_and_shards_over_ranks():
    files = [f"{i:02d}.png" for i in range(11)]
    b = infer.plan_batches(files[::-1], ["03.png", "07.png"], 4)
    assert b == [["00.png", "01.png", "02.png", "04.png"], ["05.png", "06.png", "08.png", "09.png"], ["10.png"]]
    parts = [infer.plan_batches(files, [], 2, r, 3) for r in range(3)]
    assert sorted(f for p in parts for bt in p for f in bt) == files and all(len(bt) <= 2 for p in parts for bt in p)


def test_checkpoint_key_conversion():
    sd = {"unet.down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora.down.weight": torch.zeros(64, 640),
          "unet.down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora.up.weight": torch.zeros(640, 64),
          "unet.down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k_ip.lora_A.weight": torch.zeros(64, 2048),
          "text_encoder.x.lora.down.weight": torch.zeros(1)}
    out = infer.convert_previewer_lora(sd)
    assert set(out) == {"down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora_A.weight",
                        "down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora_B.weight",
                        "down_blocks.1.attentions.0.transformer_blocks.0.attn2.processor.to_k_ip.lora_A.weight"}
    legacy = {"image_proj_model.latents": torch.zeros(1), "adapter_modules.1.to_k_ip.weight": torch.zeros(1)}
    rv = infer.revise_adapter_state_dict(legacy)
    assert set(rv["image_proj"]) == {"latents"} and set(rv["ip_adapter"]) == {"1.to_k_ip.weight"}
    ids = infer.hashed_token_ids(["a photo of a cat", ""])
    assert ids.shape == (2, 77) and int(ids[0].argmax()) == 6 and int(ids[1].argmax()) == 1 and torch.equal(ids, infer.hashed_token_ids(["a photo of a cat", ""]))


@pytest.mark.gpu
def test_cli_end_to_end_random_init(tmp_path):
    """`python -m instantir_b200.infer --random_init`: two inputs of different aspect, 2 steps, files written at the input
    sizes; a second run skips them (infer.py:151-162)."""
    import numpy as np
    from PIL import Image

    src, out = tmp_path / "lq", tmp_path / "out"
    src.mkdir()
    rng = np.random.default_rng(0)
    for name, (w, h) in (("a.png", (300, 200)), ("b.png", (256, 256))):
        Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(src / name)
    cmd = [sys.executable, "-m", "instantir_b200.infer", "--random_init", "--test_path", str(src), "--out_path", str(out),
           "--num_inference_steps", "2", "--batch_size", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    assert "restored 2 image(s)" in r.stdout
    for name, size in (("a.png", (300, 200)), ("b.png", (256, 256))):
        img = Image.open(out / name)
        assert img.size == size and np.isfinite(np.asarray(img, dtype=np.float32)).all()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "Skip a.png" in r.stdout and "restored 0 image(s)" in r.stdout
