"""Multi-GPU partitioning of the denoising loop (one process per GPU, torch.distributed plumbing).

The path shards two ways and needs exactly one collective:
  * data-parallel over images — no communication at all;
  * CFG-parallel — the uncond and cond halves of every forward are independent until
    ``noise_pred.chunk(2)`` (pipelines/sdxl_instantir.py:1620): rank 2k runs the uncond branch, rank
    2k+1 the cond branch, and ONE all-gather of eps [B,4,h,w] per step (128 KiB/img fp32 at 1024²)
    lets both ranks run the identical fused CFG+DDPM kernel with identical noise.
Backend: NCCL over NVLink/NVSwitch on GPUs; gloo on CPU for the host-logic tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def partition(n_images: int, world_size: int, rank: int, cfg_parallel: bool) -> Tuple[slice, Optional[int]]:
    """(slice of the image batch this rank restores, CFG branch or None).

    With cfg_parallel, ranks (2k, 2k+1) form a pair working on the same images (branch 0 = uncond,
    1 = cond) and the images are split across the world_size/2 pairs; otherwise across all ranks.
    Images are split contiguously, the first (n mod groups) groups taking one extra image."""
    if cfg_parallel:
        if world_size % 2:
            raise ValueError("CFG-parallel needs an even number of ranks")
        groups, g, branch = world_size // 2, rank // 2, rank % 2
    else:
        groups, g, branch = world_size, rank, None
    base, extra = divmod(n_images, groups)
    start = g * base + min(g, extra)
    return slice(start, start + base + (1 if g < extra else 0)), branch


class CFGParallel:
    """One CFG pair: holds the 2-rank process group and gathers the two eps branches."""

    def __init__(self, rank: Optional[int] = None, world_size: Optional[int] = None):
        self.rank = dist.get_rank() if rank is None else rank
        self.world_size = dist.get_world_size() if world_size is None else world_size
        if self.world_size % 2:
            raise ValueError("CFG-parallel needs an even number of ranks")
        self.branch = self.rank % 2
        self.group = None
        # every rank must take part in the creation of every pair group
        for k in range(self.world_size // 2):
            g = dist.new_group(ranks=[2 * k, 2 * k + 1])
            if k == self.rank // 2:
                self.group = g

    def broadcast_from_leader(self, t: torch.Tensor) -> torch.Tensor:
        """the uncond rank's tensor on both ranks of the pair (initial latents, DDPM noise: the two branches must
        step identical x_t with identical z, whatever RNG state each process came with)"""
        t = t.contiguous()
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else (self.rank // 2) * 2,
                       group=self.group)
        return t

    def gather_branches(self, eps: torch.Tensor) -> torch.Tensor:
        """[B,4,h,w] of this rank's branch -> [2B,4,h,w] = [uncond; cond] on both ranks of the pair."""
        eps = eps.contiguous()
        out = torch.empty((2 * eps.shape[0],) + tuple(eps.shape[1:]), device=eps.device, dtype=eps.dtype)
        chunks = [out[: eps.shape[0]], out[eps.shape[0]:]]
        dist.all_gather(chunks, eps, group=self.group)
        return out


def draw_shared_noise(shape, generator, device, sl: slice):
    """DP ranks draw the FULL-batch noise the single-GPU run would draw and keep their slice, so the
    sharded run is bit-comparable with the unsharded one (SURVEY §8e 'What does not shard')."""
    full = torch.randn(shape, generator=generator, dtype=torch.float32)
    return full[sl].to(device)
