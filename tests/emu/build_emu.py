"""TEST INFRASTRUCTURE: build tests/_emu_build/libcdbg_emu.so -- the library sources compiled for the CPU emulator.

The sources under bcalm_b200/csrc are copied unchanged except for one mechanical rewrite: the declaration of dynamic shared
memory (`extern __shared__ ... name[];`, which has no C++ spelling) becomes a pointer to the emulator's buffer.  Everything else
is handled by tests/emu/cuda_runtime.h being first on the include path.  See that header for what this is (and is not) for.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "bcalm_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_emu_build")          # two levels below the root, like csrc (relative include of include/cdbg.h)
LIB = os.path.join(OUT, "libcdbg_emu.so")
CLI = os.path.join(OUT, "bcalm_emu")                      # the unchanged command line tool linked against the emulated library

_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl", ".h")))


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in _sources()] + [os.path.join(HERE, "cuda_runtime.h"), os.path.join(HERE, "nccl.h"),
                                                          os.path.join(ROOT, "include", "cdbg.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    for f in _sources():
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        text = _DYN.sub(r"unsigned char *\1 = emu::dyn_smem();", text)
        with open(os.path.join(OUT, f), "w") as fh:
            fh.write(text)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-fno-strict-aliasing", "-w",
           "-I", HERE, "-o", LIB, os.path.join(OUT, "cdbg.cu"), "-ldl"]
    subprocess.check_call(cmd)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", CLI, os.path.join(CSRC, "bcalm_main.cpp"), "-L" + OUT, "-lcdbg_emu",
                           "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return LIB


def clean() -> None:
    shutil.rmtree(OUT, ignore_errors=True)


if __name__ == "__main__":
    print(build(force=True))
