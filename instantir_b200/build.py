"""In-tree nvcc build of the C-ABI library ``libinstantir_b200.so`` for sm_100a.

No torch extension machinery: the library has a plain C ABI (include/instantir_b200.h) and is
bound with ctypes, so it is compiled with a direct ``nvcc -shared`` call.  Objects are cached per
source by content hash so rebuilds after a one-file edit take seconds.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libinstantir_b200.so")            # 16-bit operand type: bf16
LIB_FP16 = os.path.join(HERE, "libinstantir_b200_fp16.so")  # same sources with -DIIR_FP16
SOURCES = ["api.cu", "gemm_tc.cu", "attn_tc.cu", "simt.cu", "norm.cu", "elementwise.cu", "sched.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sm_100a kernels cannot be built")


def _digest(path: str, extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for p in (path, os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "instantir_b200.h")):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _build_variant(lib_path: str, tag: str, defines, verbose, force, ptxas_info) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, f"{os.path.splitext(src)[0]}.{tag}.{_digest(sp, ' '.join(defines))}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", sp, "-o", obj]
            if ptxas_info:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(run, jobs):
                if verbose or ptxas_info or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {src} [{tag}]\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src} [{tag}]")
    stamp = os.path.join(BUILD, f"link.{tag}.stamp")
    want = "\n".join(objs)
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if force or jobs or not os.path.exists(lib_path) or have != want:
        cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        with open(stamp, "w") as f:
            f.write(want)
        keep = set(os.path.basename(o) for o in objs)
        for fn in os.listdir(BUILD):
            if fn.endswith(".o") and f".{tag}." in fn and fn not in keep:
                os.remove(os.path.join(BUILD, fn))
    return lib_path


def build(verbose: bool = False, force: bool = False, ptxas_info: bool = False) -> str:
    """Compile (if stale) both library variants; returns the path of the bf16 one."""
    _build_variant(LIB_FP16, "fp16", ["-DIIR_FP16=1"], verbose, force, False)
    return _build_variant(LIB, "bf16", [], verbose, force, ptxas_info)


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, ptxas_info="--ptxas" in sys.argv))
