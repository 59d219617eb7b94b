"""Build the CPU oracle shared library (test infrastructure; see oracle/hdp_oracle.c header).

The reference (AgentOxygen/HDP) is pure Python + Numba: it has no C/C++ sources, so there is
nothing to compile into oracle/_ref/.  The oracle is therefore the C restatement in
hdp_oracle.c, pinned against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hdp_oracle.c")
LIB = os.path.join(HERE, "libhdp_oracle.so")

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


if __name__ == "__main__":
    print(build(force=True))
