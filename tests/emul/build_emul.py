"""TEST INFRASTRUCTURE ONLY: compiles the framework's plain-CUDA sources with g++ against the CUDA execution-model
emulator (cuda_emul.h) so that kernel logic and the host-side schedules of csrc/api.cu can be checked against the
oracle on the GPU-less build container.  Kernels that use inline PTX (tcgen05/TMA: lstm_tc.cu) are NOT emulated.

The product loader (mmego_b200/_capi.py) never loads this library; only tests/test_emul_*.py do."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mmego_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmmego_emul.so")

SOURCES = ["api.cu", "gemm_ffma.cu", "point_upper.cu", "lower_frame.cu", "lstm_small.cu", "gcn.cu", "decode.cu", "snippet.cu", "heads_mma.cu",
           "lstm_resident.cu", "pack.cpp"]
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-DMMEGO_EMUL", "-I", HERE, "-I", CSRC, "-Wno-unused-result", "-Wno-attributes"]


def _digest() -> str:
    h = hashlib.sha256()
    names = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "cuda_emul.h"),
                                                                         os.path.join(HERE, "cuda_emul.cpp"),
                                                                         os.path.join(ROOT, "include", "mmego_b200.h")]
    for p in names:
        if os.path.isfile(p):
            h.update(p.encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()


def _cc(src: str) -> str:
    obj = os.path.join(OUT, os.path.basename(src) + ".o")
    cmd = ["g++", *FLAGS, "-x", "c++", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on {src}:\n{r.stderr[-4000:]}")
    return obj


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    stamp = os.path.join(OUT, "build.sha256")
    dig = _digest()
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "cuda_emul.cpp")]
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_cc, srcs))
    r = subprocess.run(["g++", "-shared", "-o", LIB, *objs, "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    open(stamp, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build())
