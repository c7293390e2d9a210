"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports exactly the
symbols include/instantir_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

from instantir_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "instantir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iir_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert _header_symbols() == sorted(_lib.SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    lib = _lib.load()
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert lib.iir_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.iir_launch_count(), int)
    assert lib.iir_groupnorm_scratch_floats(2, 32) == 2 * 256 * 32 * 2 + 2 * 32 * 2 + 2


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors vs the C compiler's own layout of the header structs."""
    import subprocess

    fields_g = ["a", "M", "lda", "conv", "bias", "rows_per_sample", "residual", "ld_res", "aux", "out",
                "ld_out", "act", "bn", "cluster", "ld_rowvec", "ln_eps", "conv_asym", "gn_sums", "gn_cpg", "gn_groups"]
    fields_a = ["q", "n_seg", "k", "ldk", "k_off", "v", "kv_len", "seg_scale", "out", "dtype", "n_q",
                "softmax_scale"]
    prog = ["#include <stdio.h>", "#include <stddef.h>", '#include "instantir_b200.h"', "int main(void){",
            'printf("%zu %zu\\n", sizeof(iir_gemm_args), sizeof(iir_attn_args));',
            'printf("%zu %zu %zu\\n", sizeof(iir_adaln_item), offsetof(iir_adaln_item, mod_off), offsetof(iir_adaln_item, C));']
    prog += [f'printf("%zu\\n", offsetof(iir_gemm_args, {f}));' for f in fields_g]
    prog += [f'printf("%zu\\n", offsetof(iir_attn_args, {f}));' for f in fields_a]
    prog += ["return 0;}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    nums = [int(x) for x in out]
    assert nums[0] == ctypes.sizeof(_lib.GemmArgs) and nums[1] == ctypes.sizeof(_lib.AttnArgs)
    assert nums[2:5] == [ctypes.sizeof(_lib.AdaLNItem), _lib.AdaLNItem.mod_off.offset, _lib.AdaLNItem.C.offset]
    nums = nums[:2] + nums[5:]
    got = [getattr(_lib.GemmArgs, f).offset for f in fields_g] + [getattr(_lib.AttnArgs, f).offset for f in fields_a]
    assert nums[2:] == got


def test_invalid_arguments_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.iir_lcm_step(None, 0, None, None, 7, 0.5, 0.0, 1.0, None)
    assert rc == -1
    assert b"iir_lcm_step" in lib.iir_last_error()
    g = _lib.GemmArgs()
    g.a_dtype = g.w_dtype = _lib.F32
    assert lib.iir_gemm_tc(ctypes.byref(g), None) == -1  # fp32 operands are refused by the tc path
    assert b"bf16" in lib.iir_last_error()


def test_opt_in_groupnorm_entry_points_validate_their_arguments():
    """ABI 11: argument checks answer with a status and a message, on a box without a GPU too"""
    lib = _lib.load()
    assert lib.iir_memset_zero(None, 16, None) == -1 and b"iir_memset_zero" in lib.iir_last_error()
    assert lib.iir_groupnorm_apply_sums(None, 0, None, None, None, None, 0, 1, 64, 64, 32, 1e-5, 0, None) == -1
    assert b"iir_groupnorm_apply_sums" in lib.iir_last_error()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf) + (-ctypes.addressof(buf)) % 256  # an aligned, non-null address: every call below is refused before a launch
    assert lib.iir_groupnorm_apply_sums(p, 0, None, None, p, p, 0, 1, 64, 100, 32, 1e-5, 0, None) == -1  # C % groups != 0
    g = _lib.GemmArgs()
    g.a = g.w = g.out = p
    g.a_dtype = g.w_dtype = g.out_dtype = lib.iir_h16_dtype()
    g.M, g.N, g.K, g.lda, g.ld_out, g.bn = 256, 320, 64, 64, 320, 160
    g.gn_sums, g.gn_groups, g.gn_cpg = p, 32, 10
    g.rows_per_sample = 100                      # a warp's 32 rows would straddle two samples
    rc = lib.iir_gemm_tc(ctypes.byref(g), None)
    assert rc != 0
    msg = lib.iir_last_error()
    assert b"rows_per_sample" in msg or b"cuTensorMapEncodeTiled" in msg or b"CUDA" in msg, msg
    g.gn_cpg = 9                                 # odd group width / N != groups * cpg
    assert lib.iir_gemm_tc(ctypes.byref(g), None) != 0
