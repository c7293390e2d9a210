"""instantir_b200 — B200-native (sm_100a) implementation of InstantIR's per-timestep denoising step.

Host code is Python/PyTorch (device memory, streams, torch.distributed plumbing); all arithmetic on
the hot path runs in the hand-written CUDA kernels of ``libinstantir_b200.so`` behind the C ABI
declared in ``include/instantir_b200.h``.  There is no CPU or torch fallback.
"""
__version__ = "0.1.0"
