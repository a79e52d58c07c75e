"""Builds ``libfus_b200.so`` (every CUDA kernel + the C ABI of
``include/fus_b200.h``) in-tree with nvcc for sm_100a.

    python -m fenicsx_fus_gpu_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting ``.so`` is git-ignored but
travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libfus_b200.so")
SOURCES = ["api.cu", "stiffness.cu", "stiffness_affine.cu", "stiffness_vertex.cu", "mass.cu", "vector.cu", "rk.cu", "geometry.cu", "halo.cu", "sampling.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("FUS_NVCC_EXTRA", "").split()  # experiments, e.g. -DFUS_CFG_ALT
# IEEE division / sqrt (no fast-math): parity with the reference is rel-L2 <= 1e-12 in f64
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-prec-div=true", "-prec-sqrt=true", "-fmad=true",
]


def _deps(src: str):
    d = [os.path.join(CSRC, src), os.path.join(CSRC, "fus_common.cuh"), os.path.join(CSRC, "stiffness_kernel.cuh"),
         os.path.join(CSRC, "halo_internal.cuh"),
         os.path.join(os.path.dirname(HERE), "include", "fus_b200.h"), os.path.abspath(__file__)]
    return [p for p in d if os.path.exists(p)]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    """True when the library is missing, or older than any source it is built from AND a
    compiler is at hand (a GPU box receives the prebuilt .so with fresh mtimes; without nvcc
    the existing library is used as is)."""
    if not os.path.exists(LIB):
        return True
    if not os.path.exists(NVCC):
        return False
    t = os.path.getmtime(LIB)
    srcs = {p for s in SOURCES for p in _deps(s)}
    return any(os.path.getmtime(p) > t for p in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, _deps(s)):
            jobs.append((s, obj))

    def compile_one(job):
        s, obj = job
        cmd = [NVCC, *FLAGS, *EXTRA, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            logs = list(ex.map(compile_one, jobs))
        if verbose:
            print("\n".join(logs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
