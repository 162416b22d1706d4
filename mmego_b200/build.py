"""Builds libmmego_b200.so (hand-written sm_100a CUDA behind the C ABI of include/mmego_b200.h) IN-TREE.

    python -m mmego_b200.build [--force] [--verbose] [--variant product|ffma]

`--variant ffma` builds the TEST-ONLY library tests/_variant_build/libmmego_b200_ffma.so from the same sources with
-DMMEGO_WITH_FFMA -DMMEGO_DEBUG_SWITCHES: it additionally holds the first-generation fp32 FFMA kernels (exact-fp32 A/B
references used by tests/test_gpu_parity.py) and the tc_dbg experiment switches.  The product library has neither.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file
(mmego_b200/lib/libmmego_b200.so), is git-ignored and travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libmmego_b200.so")

SOURCES = ["api.cu", "gemm_ffma.cu", "point_upper.cu", "lower_frame.cu", "lstm_small.cu", "gcn.cu", "decode.cu", "snippet.cu", "heads_mma.cu",
           "lstm_tc.cu", "gemm_tc.cu", "lstm_resident.cu", "pack.cpp"]
HEADERS = ["internal.h", "pack.h", "point_layout.h", "cuda_compat.h", "tc_common.cuh", "mma_frag.cuh",
           os.path.join("..", "..", "include", "mmego_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmmego_b200 cannot be built (there is no CPU fallback)")


VARIANTS = {
    "product": dict(flags=[], libdir=LIBDIR, lib=LIB),
    "ffma": dict(flags=["-DMMEGO_WITH_FFMA", "-DMMEGO_DEBUG_SWITCHES"],
                 libdir=os.path.join(ROOT, "tests", "_variant_build"),
                 lib=os.path.join(ROOT, "tests", "_variant_build", "libmmego_b200_ffma.so")),
}


def _digest(extra=()) -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        p = os.path.join(CSRC, name)
        if os.path.exists(p):
            h.update(name.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(list(NVCC_FLAGS) + list(extra)).encode())
    return h.hexdigest()


def _compile(nvcc: str, src: str, verbose: bool, objdir: str = OBJDIR, extra=()) -> str:
    obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(objdir, os.path.splitext(src)[0] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, variant: str = "product") -> str:
    v = VARIANTS[variant]
    libdir, lib, extra = v["libdir"], v["lib"], v["flags"]
    objdir = os.path.join(libdir, "obj")
    os.makedirs(objdir, exist_ok=True)
    stamp = os.path.join(libdir, "build.sha256")
    dig = _digest(extra)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return lib
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, verbose, objdir, extra), srcs))
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", default="product", choices=sorted(VARIANTS))
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.variant))


if __name__ == "__main__":
    main()
