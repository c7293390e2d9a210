"""CPU oracle for InstantIR's per-timestep denoising step — TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch fp32 restatement of the reference's hot path
(pipelines/sdxl_instantir.py:1497-1666 and the modules it drives).  It exists so that the CUDA
product in ``instantir_b200/`` can be checked against the reference's arithmetic on machines where
the reference itself cannot run (its UNet lives in the absent ``diffusers==0.28.1`` dependency).

Rules
  * Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
    reference`` legs may import it.  The product never does
    (tests/test_host_logic.py::test_product_never_imports_oracle_or_reference enforces this).
  * It never calls into ``instantir_b200``.

Pinning (how the restatement is tied to the reference)
  * The reference files that import in the authoring container were run there verbatim —
    ``module/ip_adapter/attention_processor.py`` (AttnProcessor2_0, TA_IPAttnProcessor2_0,
    AdaLayerNorm), ``module/ip_adapter/resampler.py`` (Resampler),
    ``module/ip_adapter/ip_adapter.py`` (MultiIPAdapterImageProjection),
    ``schedulers/lcm_single_step_scheduler.py`` (behind a small ``diffusers`` stub),
    ``module/min_sdxl.py`` block classes (ResnetBlock2D, Transformer2DModel, down/up/mid blocks,
    full UNet forward) and ``module/aggregator.py`` (Aggregator.forward / SFT, behind a stub whose
    diffusers blocks are this oracle's blocks) — on seeded inputs; their outputs are committed
    under ``tests/golden/`` together with ``tests/golden/make_golden.py``.
    ``tests/test_oracle_golden.py`` replays the oracle against those vectors.  Likewise the vendored VAE
    (``module/diffusers_vae/vae.py`` Decoder / Encoder / DiagonalGaussianDistribution, ``tests/golden/make_golden_vae.py``),
    ``rescale_noise_cfg`` and ``infer.py``'s ``resize_img`` / argument parser (``tests/golden/make_golden_misc.py``), and
    ``transformers``' own CLIPTextModel / CLIPTextModelWithProjection / Dinov2Model for ``oracle/encoders.py``
    (``tests/golden/make_golden_encoders.py``).
  * Arithmetic that lives only in absent third-party code (diffusers DDPMScheduler.step,
    residual injection in UNet2DConditionModel.forward, peft LoRA) is restated from the published
    algorithm and is **unpinned** by any reference-run vector; it is checked
    by closed-form / invariant tests only.
"""
