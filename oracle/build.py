"""Build the CPU oracle (test infrastructure; see oracle/hdp_oracle.c header) and materialise oracle/_ref/.

* ``build()``      compiles the C restatement hdp_oracle.c -> libhdp_oracle.so (the checker of the parity tests).
* ``build_ref()``  installs the UNMODIFIED reference package (AgentOxygen/HDP, pure Python + Numba) from
  ``/root/reference`` into ``oracle/_ref/`` with pip (offline, ``--no-deps``: its xarray/dask/cftime dependencies
  are not installable in this image and are stubbed at import time by oracle/ref_numba.py).  ``oracle/_ref/`` is
  git-ignored build output - no reference source enters the history - but it is not gpurun-ignored, so the
  reference's own Numba kernels travel to the GPU box with the built ``.so`` files and bench.py can time them
  there as the CPU baseline (``cpu_baseline.kind = "reference"``).  ``/root/reference`` itself exists only in the
  build container; when it is absent ``build_ref()`` keeps whatever ``oracle/_ref/`` already holds.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hdp_oracle.c")
LIB = os.path.join(HERE, "libhdp_oracle.so")

REF_SRC = os.environ.get("HDP_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")

_FLAGS = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden", "-Wall", "-Wextra"]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    errors = []
    # $CC in this image points at a gcc without libgomp.spec; prefer the system gcc.
    for cc in ("/usr/bin/gcc", shutil.which("gcc"), os.environ.get("CC"), shutil.which("cc")):
        if not cc:
            continue
        for omp in (["-fopenmp"], []):
            cmd = [cc, *_FLAGS, *omp, "-o", LIB + ".tmp", SRC, "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                os.replace(LIB + ".tmp", LIB)
                return LIB
            errors.append(" ".join(cmd) + "\n" + r.stderr)
    raise RuntimeError("could not build the oracle:\n" + "\n".join(errors))


def ref_installed() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "hdp", "threshold.py")) and os.path.isfile(os.path.join(REF_DIR, "hdp", "metric.py"))


def build_ref(force: bool = False):
    """``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>``.
    Returns the directory, or None when there is neither a reference tree nor an earlier install."""
    if not os.path.isdir(os.path.join(REF_SRC, "hdp")):
        return REF_DIR if ref_installed() else None
    if ref_installed() and not force:
        newest = max(os.path.getmtime(os.path.join(REF_SRC, "hdp", f)) for f in os.listdir(os.path.join(REF_SRC, "hdp")) if f.endswith(".py"))
        if os.path.getmtime(os.path.join(REF_DIR, "hdp", "metric.py")) >= newest:
            return REF_DIR
    with tempfile.TemporaryDirectory(prefix="hdp_ref_src_") as tmp:
        src = os.path.join(tmp, "reference")              # the build writes egg-info into the tree: /root/reference is read-only
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "imgs", "__pycache__"))
        stage = os.path.join(tmp, "target")
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", stage, src]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        if r.returncode != 0:
            raise RuntimeError("could not install the reference into oracle/_ref:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        shutil.rmtree(REF_DIR, ignore_errors=True)
        shutil.copytree(stage, REF_DIR, ignore=shutil.ignore_patterns("__pycache__"))
    return REF_DIR


if __name__ == "__main__":
    print(build(force=True))
    print(build_ref(force=True))
