"""Oracle LoRA (peft==0.10.0 semantics): y = base(x) + (alpha/r) * B(A(x)).

peft is absent in this environment: **parity unpinned**; checked by invariants (B = 0 => base,
merged-weight equivalence).  Target matching follows peft: a module is wrapped when its dotted
name equals a target or ends with "." + target (targets: pipelines/sdxl_instantir.py:141-162).
"""
from __future__ import annotations

import torch
import torch.nn as nn

PREVIEWER_LORA_MODULES = [
    "to_q", "to_kv", "0.to_out", "attn1.to_k", "attn1.to_v", "to_k_ip", "to_v_ip", "ln_k_ip.linear",
    "ln_v_ip.linear", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2", "conv1", "conv2",
    "conv_shortcut", "downsamplers.0.conv", "upsamplers.0.conv", "time_emb_proj",
]


class _Switch:
    def __init__(self):
        self.enabled = False


class LoRALinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float, switch: _Switch):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Linear(base.in_features, r, bias=False)
        self.lora_B = nn.Linear(r, base.out_features, bias=False)
        nn.init.zeros_(self.lora_B.weight)
        self.scaling = alpha / r
        self._switch = switch

    @property
    def weight(self):  # lets code that reads `.weight` (e.g. init_attn_proc) keep working
        return self.base_layer.weight

    @property
    def in_features(self):
        return self.base_layer.in_features

    def forward(self, x):
        y = self.base_layer(x)
        if self._switch.enabled:
            y = y + self.lora_B(self.lora_A(x)) * self.scaling
        return y


class LoRAConv2d(nn.Module):
    def __init__(self, base: nn.Conv2d, r: int, alpha: float, switch: _Switch):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride, base.padding, bias=False)
        self.lora_B = nn.Conv2d(r, base.out_channels, 1, bias=False)
        nn.init.zeros_(self.lora_B.weight)
        self.scaling = alpha / r
        self._switch = switch

    def forward(self, x):
        y = self.base_layer(x)
        if self._switch.enabled:
            y = y + self.lora_B(self.lora_A(x)) * self.scaling
        return y


def _matches(name: str, targets) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def add_previewer_lora(unet: nn.Module, r: int, alpha: float, targets=PREVIEWER_LORA_MODULES):
    """unet.add_adapter(LoraConfig(r, targets, alpha)) + disable_adapters()
    (pipelines/sdxl_instantir.py:376-395).  Adds enable_adapters()/disable_adapters() to `unet`."""
    switch = _Switch()
    todo = []
    for name, m in unet.named_modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)) and _matches(name, targets):
            todo.append(name)
    for name in todo:
        parent_name, _, attr = name.rpartition(".")
        parent = unet.get_submodule(parent_name) if parent_name else unet
        m = getattr(parent, attr) if not attr.isdigit() else parent[int(attr)]
        wrapped = LoRALinear(m, r, alpha, switch) if isinstance(m, nn.Linear) else LoRAConv2d(m, r, alpha, switch)
        if attr.isdigit():
            parent[int(attr)] = wrapped
        else:
            setattr(parent, attr, wrapped)
    unet._lora_switch = switch
    unet.enable_adapters = lambda: setattr(switch, "enabled", True)
    unet.disable_adapters = lambda: setattr(switch, "enabled", False)
    return todo


def merged_weight(mod) -> torch.Tensor:
    """W + (alpha/r) * B·A — the second weight set the CUDA product uses for the previewer pass."""
    w = mod.base_layer.weight.data
    if isinstance(mod, LoRALinear):
        return w + mod.scaling * (mod.lora_B.weight.data @ mod.lora_A.weight.data)
    a = mod.lora_A.weight.data  # [r, Cin, k, k]
    b = mod.lora_B.weight.data[:, :, 0, 0]  # [Cout, r]
    return w + mod.scaling * torch.einsum("or,rikl->oikl", b, a)
