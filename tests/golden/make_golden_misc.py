"""Generate tests/golden/rescale_noise_cfg.pt by executing the REFERENCE's own ``rescale_noise_cfg``
(pipelines/sdxl_instantir.py:179-192) — only that function's source is exec'd, because the pipeline file imports
diffusers at module level.  Needs /root/reference (authoring container only).

    python tests/golden/make_golden_misc.py [--ref /root/reference]
"""
import argparse
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from seeding import rnd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    src = open(os.path.join(args.ref, "pipelines", "sdxl_instantir.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def rescale_noise_cfg"))
    end = next(i for i in range(start + 1, len(src)) if src[i] and not src[i].startswith((" ", "\t")))
    ns = {"torch": torch}
    exec("\n".join(src[start:end]), ns)
    fn = ns["rescale_noise_cfg"]
    e_u, e_c = rnd(3, 4, 16, 24, seed=91), rnd(3, 4, 16, 24, seed=92) * 1.4 + 0.1
    cfg = e_u + 7.0 * (e_c - e_u)
    out = {phi: fn(cfg, e_c, guidance_rescale=phi) for phi in (0.0, 0.3, 0.7, 1.0)}
    torch.save({"e_u": e_u, "e_c": e_c, "guidance": 7.0, "out": out}, os.path.join(HERE, "rescale_noise_cfg.pt"))
    print("rescale_noise_cfg.pt", {k: float(v.std()) for k, v in out.items()})

    # ---- infer.py: resize_img (:31-66) executed verbatim on blank PIL images, and the CLI's flag names (:229-386)
    import json
    import re

    import numpy as np
    from PIL import Image

    src = open(os.path.join(args.ref, "infer.py")).read()
    lines = src.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def resize_img"))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i].startswith((" ", "\t")))
    ns = {"Image": Image, "np": np}
    exec("\n".join(lines[start:end]), ns)
    cases = []
    for (w, h, width, height) in [(256, 256, None, None), (1024, 1024, None, None), (1500, 1000, None, None), (640, 480, None, None),
                                   (3000, 500, None, None), (513, 777, None, None), (800, 600, 1024, None), (800, 600, None, 512),
                                   (800, 600, 640, 640), (400, 1200, None, None), (1023, 769, None, None)]:
        img, out_size = ns["resize_img"](Image.new("RGB", (w, h)), width=width, height=height)
        cases.append({"w": w, "h": h, "width": width, "height": height, "runtime": list(img.size), "out": list(out_size)})
    flags = sorted(set(re.findall(r'"(--[a-z_0-9]+)"', src)))
    json.dump({"resize_img": cases, "flags": flags}, open(os.path.join(HERE, "infer_cli.json"), "w"), indent=1)
    print("infer_cli.json", len(cases), "resize cases,", len(flags), "flags")


if __name__ == "__main__":
    main()
