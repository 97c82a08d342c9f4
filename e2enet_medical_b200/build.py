"""Builds libe2enet_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m e2enet_medical_b200.build [--force]

The .so and the _build/ objects + ptxas logs are git-ignored (.gitignore) but travel to the GPU box with the
repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libe2enet_b200.so")

SOURCES = ["api.cu", "gather_gemm.cu", "conv_tc.cu", "elementwise.cu", "masking.cu", "window.cu", "loss.cu", "optim.cu", "export.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-DE2E_B200=1", "-I" + INCLUDE,
              "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libe2enet_b200.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "e2enet_b200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return job, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for (s, o), r in ex.map(compile_one, jobs):
                log = (r.stdout or "") + (r.stderr or "")
                with open(o + ".log", "w") as f:
                    f.write(log)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s" % (s, log[-6000:]))
                if verbose:
                    print(log)
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
