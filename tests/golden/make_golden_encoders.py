"""Golden vectors for SURVEY §8 row f2, produced by the `transformers` build installed in the authoring container (run
once by hand: `python tests/golden/make_golden_encoders.py`; the .pt it writes is committed):

  * CLIPTextModel (quick_gelu, CLIP-L style) and CLIPTextModelWithProjection (gelu, bigG style) on seeded token ids:
    hidden_states[-2], last_hidden_state / text_embeds, pooler_output — what encode_prompt reads
    (/root/reference/pipelines/sdxl_instantir.py:522-533);
  * Dinov2Model on a seeded image at the table's own resolution and at a smaller one (interpolated position
    embeddings): last_hidden_state — what encode_image reads (:659-667).

Small random-init configurations with head_dim 64 (the product's attention kernel); weights are re-created from the
seed by tests/golden/seeding.py, so only seeds, inputs and outputs are stored."""
import os
import sys

import torch
import transformers
from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection, Dinov2Config, Dinov2Model

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from seeding import checksum, seeded_init  # noqa: E402

torch.set_grad_enabled(False)
out = {"transformers_version": transformers.__version__}
g = torch.Generator().manual_seed(5)
ids = torch.randint(3, 990, (2, 77), generator=g)
ids[0, 20:] = 0
ids[0, 19] = 999   # the highest id marks the EOS position (legacy eos_token_id == 2 rule: argmax)
ids[1, 76] = 999
out["ids"] = ids
for name, cls, kw in (("clip_l", CLIPTextModel, dict(hidden_act="quick_gelu")),
                      ("clip_g", CLIPTextModelWithProjection, dict(hidden_act="gelu", projection_dim=96))):
    cfg = dict(vocab_size=1000, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
               max_position_embeddings=77, layer_norm_eps=1e-5, eos_token_id=2, bos_token_id=0, pad_token_id=1, **kw)
    m = seeded_init(cls(CLIPTextConfig(**cfg)).eval(), 41 if name == "clip_l" else 42)
    r = m(ids, output_hidden_states=True)
    rec = dict(cfg={k: v for k, v in cfg.items() if k not in ("bos_token_id", "pad_token_id")}, seed=41 if name == "clip_l" else 42,
               checksum=checksum(m), names=sorted(k for k, _ in m.named_parameters()),
               penultimate=r.hidden_states[-2].clone(), n_hidden=len(r.hidden_states))
    if name == "clip_l":
        rec.update(last_hidden_state=r.last_hidden_state.clone(), pooler_output=r.pooler_output.clone())
    else:
        rec.update(text_embeds=r.text_embeds.clone(), last_hidden_state=r.last_hidden_state.clone())
    out[name] = rec
dcfg = dict(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, mlp_ratio=4, image_size=70, patch_size=14, num_channels=3,
            layer_norm_eps=1e-6)
dm = seeded_init(Dinov2Model(Dinov2Config(**dcfg)).eval(), 43)
x70 = torch.randn(2, 3, 70, 70, generator=g)
x42 = torch.randn(1, 3, 42, 56, generator=g)
out["dinov2"] = dict(cfg=dcfg, seed=43, checksum=checksum(dm), names=sorted(k for k, _ in dm.named_parameters()),
                     x70=x70, out70=dm(x70).last_hidden_state.clone(), x42=x42, out42=dm(x42).last_hidden_state.clone())
torch.save(out, os.path.join(HERE, "encoders.pt"))
print({k: (v if not isinstance(v, dict) else list(v)) for k, v in out.items()})
