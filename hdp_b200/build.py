"""Build libhdp_b200.so (hand-written sm_100a CUDA + the C ABI of include/hdp_b200.h) in-tree.

    python -m hdp_b200.build [--force]

The library is compiled for sm_100a only; there is no other back end and no CPU fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libhdp_b200.so")
SOURCES = ["abi.cu", "threshold.cu", "metric.cu", "measure.cu", "host.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",                 # the f64 interpolation must stay separately rounded (bit-exact parity)
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=default",
    "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhdp_b200.so cannot be built")


def sources() -> list:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(INCLUDE, "hdp_b200.h"), __file__]
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-o", LIB + ".tmp", *sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
