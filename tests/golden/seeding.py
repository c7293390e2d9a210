"""Deterministic parameter initialisation shared by make_golden.py (reference modules) and the
tests (oracle / product modules): parameters are visited in sorted-name order and drawn from one
seeded CPU generator, so two modules with identical parameter names and shapes get identical
weights.  Matrices ~ N(0, 1/fan_in); vectors ~ N(0, 0.05) (+1 for norm gains).  Nothing stays
zero-initialised, so zero-conv / adaLN / LoRA-B paths are visible to parity checks."""
import torch


def seeded_init(module, seed: int, std: float = 0.05):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
            if p.ndim >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * (fan_in ** -0.5))
            else:
                is_norm_gain = "norm" in name and name.endswith("weight")
                p.copy_(torch.randn(p.shape, generator=g) * std + (1.0 if is_norm_gain else 0.0))
    return module


def checksum(module) -> float:
    """order-independent fingerprint of a module's parameters (float64 sum of |p| and count)."""
    tot = 0.0
    for _, p in module.named_parameters():
        tot += float(p.detach().double().abs().sum())
    return tot


def rnd(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
