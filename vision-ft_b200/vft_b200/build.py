"""Compile csrc/*.cu into vft_b200/libvft_b200.so for sm_100a (in-tree; the .so travels with gpurun).

nvcc cross-compiles without a GPU, so this is also the "does it build" check of
``__graft_entry__.build()``.  Every source is compiled to its own object (in parallel, only when it or a header is
newer than the object) and the objects are linked into the shared library.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(HERE, "libvft_b200.so")
SOURCES = ["vft_api.cu", "nf4_quant.cu", "absmax_nest.cu", "qlora_simt.cu", "qlora_gemv.cu", "lora_mma.cu", "lora_tc.cu",
           "qlora_tc.cu", "qlora_tc2.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-ccbin", "g++",
]


def _headers() -> list[str]:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "vft_b200.h"))
    return deps


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose and proc.stderr:
        print(proc.stderr, file=sys.stderr)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src, *hdrs]):
            jobs.append([nvcc, *NVCC_FLAGS, *(extra_flags or []), "-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda c: _run(c, verbose), jobs))
    if jobs or _newer(LIB, objs):
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-ccbin", "g++", "-o", LIB, *objs], verbose)
    return LIB


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose=True, extra_flags=flags))
