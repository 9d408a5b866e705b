"""Build libscn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python occuseg_b200/csrc/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["capi.cu", "meta.cu", "io.cu", "conv_simt.cu", "conv_small.cu", "conv_tma.cu", "bn.cu", "scatter.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "scn_b200.h")]
LIB = os.path.join(HERE, "libscn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("SCN_NVCC_EXTRA", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append([NVCC] + FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        return r.stderr

    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            logs = list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB, "\n".join(logs)


if __name__ == "__main__":
    lib, log = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    if "--ptxas" in sys.argv:
        print(log)
    print(lib)
