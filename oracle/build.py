"""Compile the C part of the oracle (gcc, no contraction so fp32 comparisons are reproducible)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libwd_oracle.so")
SRC = [os.path.join(HERE, "postprocess_ref.c")]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = SRC + [os.path.join(HERE, "..", "include", "wedetect_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB] + SRC + ["-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
