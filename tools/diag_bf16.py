"""bf16 error budget on config 1: eps error per module vs the fp32 oracle, and how the DDPM step scales it."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import build_oracle, export_state, make_inputs, rel_l2
from instantir_b200 import config as pcfg, weights
from instantir_b200.aggregator import Aggregator
from instantir_b200.unet import UNet2DConditionModel
from oracle import config as ocfg, schedulers as osched
torch.set_grad_enabled(False)
DEV = "cuda"
oc = ocfg.tiny(); alpha = 8.0
ounet, oagg = build_oracle(oc, 0, alpha)
if os.environ.get("IIR_DIAG_ROUND_W") == "1":  # weights as a 16-bit checkpoint would hold them: exactly bf16-representable
    for m in (ounet, oagg):
        for p_ in m.parameters():
            p_.data = p_.data.to(torch.bfloat16).float()
    print("oracle weights rounded to bf16-representable values")
inp = make_inputs(oc)
usd, ulora = export_state(ounet); asd, _ = export_state(oagg)
pc = pcfg.ModelConfig(**oc.to_dict())
x = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(5))
text = torch.cat([inp["negative_prompt_embeds"], inp["prompt_embeds"]]); pooled = torch.cat([inp["negative_pooled_prompt_embeds"], inp["pooled_prompt_embeds"]])
tid = inp["time_ids"].repeat(2, 1); ip = [torch.cat([inp["ip"][0], inp["ip"][1]]).unsqueeze(1)]
img = torch.cat([inp["image"]] * 2)
for t in (958, 501, 34):
    added = {"text_embeds": pooled, "time_ids": tid, "image_embeds": ip}
    emb = ounet.time_embedding(ounet.get_time_embed(x, torch.tensor(t)))
    emb = emb + ounet.get_aug_embed(emb, text, added)
    ck = {"temb": emb}
    ref_e = ounet(x, torch.tensor(t), text, added_cond_kwargs=added, cross_attention_kwargs=ck)[0]
    od, om_ = oagg(img, torch.tensor(t), text, controlnet_cond=x, added_cond_kwargs=added)
    ref_er = ounet(x, torch.tensor(t), text, added_cond_kwargs=added, cross_attention_kwargs=ck, down_block_additional_residuals=od, mid_block_additional_residual=om_)[0]
    for prec in ("bf16", "fp16", "fp32"):
        unet = UNet2DConditionModel(pc, weights.StateDictSource(usd, DEV, lora=ulora, lora_scale=alpha / oc.lora_rank), DEV, prec)
        agg = Aggregator(pc, weights.StateDictSource(asd, DEV), DEV, prec)
        addd = {"text_embeds": pooled.to(DEV), "time_ids": tid.to(DEV), "image_embeds": [ip[0].to(DEV)]}
        e = unet(x.to(DEV), torch.tensor(t), text.to(DEV), added_cond_kwargs=addd)[0]
        d, m = agg(img.to(DEV), torch.tensor(t), text.to(DEV), controlnet_cond=x.to(DEV), added_cond_kwargs=addd)
        er = unet(x.to(DEV), torch.tensor(t), text.to(DEV), added_cond_kwargs=addd, down_block_additional_residuals=d, mid_block_additional_residual=m)[0]
        errs = [rel_l2(a, b) for a, b in zip(d, od)] + [rel_l2(m, om_)]
        g = 7.0
        cfg_o = ref_er[:1] + g * (ref_er[1:] - ref_er[:1]); cfg_p = er[:1].cpu() + g * (er[1:].cpu() - er[:1].cpu())
        print(f"t={t} {prec}: eps(no res) {rel_l2(e, ref_e):.2e}  eps(with res) {rel_l2(er, ref_er):.2e}  cfg-eps {rel_l2(cfg_p, cfg_o):.2e} "
              f"agg res max {max(errs):.2e} mean {sum(errs)/len(errs):.2e}  |eps|/|x| {float(ref_er.norm()/x.norm()):.2f} |cfg|/|x| {float(cfg_o.norm()/x[:1].norm()):.2f}")
s = osched.DDPMScheduler()
for n in (2, 30):
    s.set_timesteps(n)
    for t in s.timesteps[:2]:
        a, c0, c1, sig = s.coefficients(t)
        print(f"steps={n} t={int(t)}: d(prev)/d(eps) = {float(c0 * (1 - a) ** 0.5 / a ** 0.5):.3f}")
