"""Builds libmmego_b200.so (hand-written sm_100a CUDA behind the C ABI of include/mmego_b200.h) IN-TREE.

    python -m mmego_b200.build [--force] [--verbose]

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
           "lstm_tc.cu", "gemm_tc.cu", "pack.cpp"]
HEADERS = ["internal.h", "pack.h", "point_layout.h", "cuda_compat.h", "tc_common.cuh", "mma_frag.cuh",
           os.path.join("..", "..", "include", "mmego_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmmego_b200 cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        p = os.path.join(CSRC, name)
        if os.path.exists(p):
            h.update(name.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(nvcc: str, src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, verbose), srcs))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))


if __name__ == "__main__":
    main()
