"""Build libwedetect_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwedetect_b200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_split.cu", "rowops.cu", "postprocess.cu", "preprocess.cu", "mlp_fused.cu", "jpeg.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# exact-arithmetic files: no FMA contraction so the CPU oracle can reproduce every comparison
PER_FILE = {"postprocess.cu": ["-fmad=false"]}


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None, tag=""):
    """defines / out / tag: A/B variants of the library (e.g. defines=("-DWD_SPLIT_EPI_BUFS=2",), out=".../lib_b.so", tag="_b")."""
    nvcc = _nvcc()
    OUT = out or globals()["OUT"]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "wedetect_b200.h"))
    objdir = os.path.join(HERE, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, "-c", s, "-o", o] + ARCH + COMMON + PER_FILE.get(src, []) + list(defines)
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r
    with ThreadPoolExecutor(max_workers=4) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for " + cmd[2])
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ARCH + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
