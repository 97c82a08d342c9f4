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
# the same sources with -DE2E_FP16: fp16 activations / gradients / packed weights (the reference's shipped AMP
# arithmetic) instead of bf16; selected at run time with E2E_PRECISION=fp16 or _lib.set_precision("fp16")
VARIANTS = {"bf16": ("libe2enet_b200.so", "", []), "fp16": ("libe2enet_b200_fp16.so", ".fp16", ["-DE2E_FP16=1"])}

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


def build_library(force: bool = False, verbose: bool = False, variants=("bf16", "fp16")) -> str:
    """compiles every variant (objects of all variants in one parallel pass); returns the default library's path"""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "e2enet_b200.h"))
    objs, jobs = {v: [] for v in variants}, []
    for v in variants:
        _, tag, defs = VARIANTS[v]
        for src in SOURCES:
            s = os.path.join(CSRC, src)
            o = os.path.join(BUILD, src.replace(".cu", tag + ".o"))
            objs[v].append(o)
            if force or _stale(o, [s] + headers):
                jobs.append((s, o, defs))

    def compile_one(job):
        s, o, defs = job
        cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return job, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(jobs))) as ex:
            for (s, o, _), r in ex.map(compile_one, jobs):
                log = (r.stdout or "") + (r.stderr or "")
                with open(o + ".log", "w") as f:
                    f.write(log)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s" % (s, log[-6000:]))
                if verbose:
                    print(log)
    for v in variants:
        lib = os.path.join(PKG, VARIANTS[v][0])
        if force or _stale(lib, objs[v]):
            cmd = [nvcc, "-shared", "-o", lib] + objs[v] + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
