"""ctypes binding of ``libinstantir_b200.so`` (the C ABI declared in include/instantir_b200.h).

The product path has no CPU or torch fallback: if the library is missing, or a call fails, an
exception is raised.  ``load()`` builds the library in-tree with nvcc when it is absent.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libinstantir_b200.so")            # 16-bit operands: bf16
LIB_PATH_FP16 = os.path.join(HERE, "libinstantir_b200_fp16.so")  # 16-bit operands: fp16

ABI_VERSION = 11
F32, BF16, F16 = 0, 1, 2
# 16-bit operand type of the default library build: IEEE fp16 is the reference's own inference precision
# (infer.py:119) and the one that meets the north star's <= 1e-2 per-step latent bar; the bf16
# build stays available as precision="bf16"
DEFAULT_H16 = F16
ACT_NONE, ACT_SILU, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2, 3
PAIR_NONE, PAIR_GEGLU, PAIR_SFT = 0, 1, 2

# every symbol include/instantir_b200.h declares (tests/test_abi.py checks the two lists agree)
SYMBOLS = [
    "iir_abi_version", "iir_h16_dtype", "iir_last_error", "iir_launch_count",
    "iir_gemm_tc", "iir_gemm_simt", "iir_conv3x3_direct",
    "iir_attn_workspace_bytes", "iir_attn_tc", "iir_attn_simt",
    "iir_groupnorm_scratch_floats", "iir_groupnorm", "iir_groupnorm_apply_sums", "iir_memset_zero", "iir_layernorm", "iir_adaln_batched", "iir_softmax_rows",
    "iir_concat_inject", "iir_upsample2x", "iir_im2col3x3_s2", "iir_cast2d", "iir_silu", "iir_add", "iir_scale",
    "iir_timestep_embedding", "iir_linear_small", "iir_embed_tokens", "iir_patchify", "iir_vit_assemble",
    "iir_step_prologue", "iir_adastep_update", "iir_lcm_step", "iir_cfg_ddpm_step", "iir_cfg_rescale", "iir_add_noise", "iir_gaussian_sample",
]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p),
        ("a_dtype", C.c_int), ("w_dtype", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_int64),
        ("conv", C.c_int),
        ("n_img", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
        ("stride", C.c_int), ("up2", C.c_int),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p),
        ("rows_per_sample", C.c_int),
        ("residual", C.c_void_p), ("res_dtype", C.c_int), ("ld_res", C.c_int64),
        ("aux", C.c_void_p), ("aux_dtype", C.c_int), ("ld_aux", C.c_int64),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("ld_out", C.c_int64),
        ("act", C.c_int), ("pair", C.c_int), ("bn", C.c_int), ("cluster", C.c_int),
        ("ld_rowvec", C.c_int64),
        ("ln_stats_out", C.c_void_p), ("ln_out16", C.c_void_p), ("ld_ln_out16", C.c_int64),
        ("ln_stats_in", C.c_void_p), ("ln_stats_zero", C.c_void_p), ("ln_colsum", C.c_void_p), ("ln_eps", C.c_float),
        ("conv_asym", C.c_int),
        ("gn_sums", C.c_void_p), ("gn_cpg", C.c_int), ("gn_groups", C.c_int),
    ]


class AdaLNItem(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("mod_off", C.c_int64), ("C", C.c_int), ("pad_", C.c_int)]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64), ("q_off", C.c_int),
        ("n_seg", C.c_int),
        ("k", C.c_void_p * 2), ("ldk", C.c_int64 * 2), ("k_off", C.c_int * 2),
        ("v", C.c_void_p * 2), ("ldv", C.c_int64 * 2), ("v_off", C.c_int * 2),
        ("kv_len", C.c_int * 2),
        ("seg_scale", C.c_float * 2),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_off", C.c_int),
        ("dtype", C.c_int),
        ("B", C.c_int), ("heads", C.c_int), ("n_q", C.c_int),
        ("softmax_scale", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("causal", C.c_int),
    ]


class IIRError(RuntimeError):
    pass


_libs = {}
_lock = threading.Lock()


def _declare(lib):
    vp, i, i64, f = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.iir_abi_version.restype = i
    lib.iir_h16_dtype.restype = i
    lib.iir_last_error.restype = C.c_char_p
    lib.iir_launch_count.restype = C.c_uint64
    lib.iir_gemm_tc.argtypes = [C.POINTER(GemmArgs), vp]
    lib.iir_gemm_simt.argtypes = [C.POINTER(GemmArgs), vp]
    lib.iir_conv3x3_direct.argtypes = [vp, i, i, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]
    lib.iir_attn_workspace_bytes.argtypes = [i, i, i]
    lib.iir_attn_workspace_bytes.restype = i64
    lib.iir_attn_tc.argtypes = [C.POINTER(AttnArgs), vp]
    lib.iir_attn_simt.argtypes = [C.POINTER(AttnArgs), vp]
    lib.iir_groupnorm_scratch_floats.argtypes = [i, i]
    lib.iir_groupnorm_scratch_floats.restype = i64
    lib.iir_groupnorm.argtypes = [vp, i, vp, vp, vp, i, i, i, i, i, f, i, vp, vp]
    lib.iir_groupnorm_apply_sums.argtypes = [vp, i, vp, vp, vp, vp, i, i, i, i, i, f, i, vp]
    lib.iir_memset_zero.argtypes = [vp, i64, vp]
    lib.iir_layernorm.argtypes = [vp, i, vp, vp, vp, i64, i, vp, i, i, i, f, vp]
    lib.iir_adaln_batched.argtypes = [vp, i, i, i, vp, i64, f, i, vp]
    lib.iir_softmax_rows.argtypes = [vp, i64, vp, i, i64, i, i, f, vp]
    lib.iir_concat_inject.argtypes = [vp, i, i, vp, i, vp, i, i, vp, i, vp, i, vp, i, i64, vp]
    lib.iir_upsample2x.argtypes = [vp, i, vp, i, i, i, i, i, vp]
    lib.iir_im2col3x3_s2.argtypes = [vp, i, vp, i, i, i, i, i, i, vp]
    lib.iir_cast2d.argtypes = [vp, i, i64, vp, i, i64, i64, i, vp]
    lib.iir_silu.argtypes = [vp, i, vp, i, i64, vp]
    lib.iir_add.argtypes = [vp, i, vp, i, vp, i, i64, vp]
    lib.iir_scale.argtypes = [vp, i, vp, i, i64, f, vp]
    lib.iir_step_prologue.argtypes = [vp, i64, i, vp, f, vp, f, vp, i, vp]
    lib.iir_timestep_embedding.argtypes = [vp, i, i, vp, i, vp]
    lib.iir_linear_small.argtypes = [vp, i, vp, i, vp, vp, i, i, i, i, i, vp]
    lib.iir_embed_tokens.argtypes = [vp, i, i, vp, i, vp, i, vp, vp]
    lib.iir_patchify.argtypes = [vp, i, i, i, i, i, vp, i, i, vp]
    lib.iir_vit_assemble.argtypes = [vp, vp, vp, vp, i, i, i, vp]
    lib.iir_adastep_update.argtypes = [vp, vp, vp, vp, vp, i, i, i64, f, f, vp]
    lib.iir_cfg_rescale.argtypes = [vp, vp, vp, i64, i64, f, f, vp]
    lib.iir_gaussian_sample.argtypes = [vp, vp, vp, i64, i64, f, vp]
    lib.iir_lcm_step.argtypes = [vp, i, vp, vp, i64, f, f, f, vp]
    lib.iir_cfg_ddpm_step.argtypes = [vp, vp, i, vp, vp, vp, vp, i64, f, f, f, f, f, vp]
    lib.iir_add_noise.argtypes = [vp, vp, vp, i64, f, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError here == missing export
        if fn.restype is C.c_int and name not in ("iir_abi_version",):
            fn.restype = C.c_int


def load(build_if_missing: bool = True, h16: int = None):
    """Return the library whose 16-bit operand type is `h16` (BF16 or F16); raises IIRError when it
    cannot be loaded (there is no fallback)."""
    if h16 is None:
        h16 = DEFAULT_H16
    with _lock:
        if h16 in _libs:
            return _libs[h16]
        path = LIB_PATH if h16 == BF16 else LIB_PATH_FP16
        if h16 == BF16 and os.environ.get("IIR_LIB_OVERRIDE"):  # measurement builds (tools/probe_epilogue.py)
            path = os.environ["IIR_LIB_OVERRIDE"]
        if not os.path.exists(path):
            if not build_if_missing:
                raise IIRError(f"{path} not found; run `python -m instantir_b200.build`")
            from . import build as _build

            _build.build()
        try:
            lib = C.CDLL(path)
        except OSError as e:  # pragma: no cover
            raise IIRError(f"cannot load {path}: {e}") from e
        _declare(lib)
        if lib.iir_abi_version() != ABI_VERSION:
            raise IIRError(f"ABI version mismatch: library {lib.iir_abi_version()}, binding {ABI_VERSION}")
        if lib.iir_h16_dtype() != h16:
            raise IIRError(f"{path} was built for 16-bit dtype {lib.iir_h16_dtype()}, expected {h16}")
        _libs[h16] = lib
        return lib


def check(rc: int, what: str = "", lib=None):
    if rc != 0:
        msg = (lib or load()).iir_last_error().decode("utf-8", "replace")
        raise IIRError(f"{what or 'iir call'} failed ({rc}): {msg}")


def launch_count() -> int:
    with _lock:
        return sum(int(lib.iir_launch_count()) for lib in _libs.values())
