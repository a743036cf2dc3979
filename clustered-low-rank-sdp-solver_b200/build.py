"""Build csrc/libclrsdp.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).

    python build.py            # incremental (per-object timestamps), parallel
nvcc cross-compiles without a GPU. No torch headers are involved: the library is plain CUDA behind a C ABI.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["gemm_i8.cu", "linalg.cu", "solver.cu", "multi.cu", "capi.cu"]
HEADERS = ["mpf.cuh", "common.cuh", "gemm_i8.cuh", "linalg.cuh", "solver.cuh", "multi.cuh", "comm.cuh", os.path.join("..", "..", "include", "clrsdp.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
FLAGS += os.environ.get("CLRSDP_EXTRA_NVCC_FLAGS", "").split()
LIB = os.path.join(CSRC, "libclrsdp.so")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=True):
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        if _stale(obj, [os.path.join(CSRC, src)] + hdrs):
            jobs.append((src, obj))

    def run(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, cwd=CSRC)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(CSRC, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    build()
    print(LIB)
