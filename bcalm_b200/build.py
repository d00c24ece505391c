"""Build the native library in-tree (bcalm_b200/libcdbg_b200.so) with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcdbg_b200.so")
CLI = os.path.join(HERE, "bcalm")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def sources(cli: bool = False):
    inc = os.path.join(os.path.dirname(HERE), "include", "cdbg.h")
    fs = [f for f in sorted(os.listdir(CSRC)) if (f == "bcalm_main.cpp") == cli]
    return [os.path.join(CSRC, f) for f in fs] + [inc]


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "cdbg.cu"), "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False) -> str | None:
    main = os.path.join(CSRC, "bcalm_main.cpp")
    if not os.path.exists(main):
        return None
    if force or _stale(CLI, sources(cli=True) + [LIB]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", CLI, main, "-L" + HERE, "-lcdbg_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return CLI


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
    print(build_cli(force=True))
