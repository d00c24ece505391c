#!/usr/bin/env python
"""Compare the code generation (registers, spills, stack) of every kernel between a git revision and the working tree.

    python scripts/ptxas_diff.py <git-rev>            # e.g. the commit whose kernels were last timed on a GPU

No GPU needed (nvcc cross-compiles for sm_100a).  Use it before a GPU call: a change that looks innocent in the source (a run-time
switch inside a hot loop, two more live values) can move a kernel from 60 registers to 40 + spills, and the emulator cannot see that.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas=-v", "-shared"]


def report(tree: str) -> dict:
    out = os.path.join(tempfile.mkdtemp(prefix="ptxas_"), "x.so")
    p = subprocess.run(["nvcc", *FLAGS, "-o", out, os.path.join(tree, "bcalm_b200", "csrc", "cdbg.cu"), "-ldl"], capture_output=True, text=True)
    if p.returncode != 0:
        sys.exit(p.stderr[-3000:])
    res, name = {}, None
    for line in p.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for", line)
        if m:
            name = m.group(1)
            res[name] = {}
            continue
        if name is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            res[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            res[name]["regs"] = int(m.group(1))
    return res


def demangle(names):
    try:
        out = subprocess.run(["c++filt", *names], capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    rev = sys.argv[1]
    old_tree = tempfile.mkdtemp(prefix="ptxas_src_")
    tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "bcalm_b200/csrc", "include"], capture_output=True, check=True).stdout
    subprocess.run(["tar", "-x", "-C", old_tree], input=tar, check=True)
    old, new = report(old_tree), report(ROOT)
    names = demangle(sorted(set(old) | set(new)))
    fmt = lambda d: "-" if d is None else f"{d.get('regs', '?'):>3} regs, stack {d.get('stack', 0):>3}, spills {d.get('spill_st', 0)}/{d.get('spill_ld', 0)}"  # noqa: E731
    changed = 0
    for n in sorted(names, key=lambda x: names[x]):
        a, b = old.get(n), new.get(n)
        if a != b:
            changed += 1
            print(f"{names[n][:90]:90s}  {fmt(a):36s} -> {fmt(b)}")
    print(f"{changed} of {len(names)} kernels differ between {rev} and the working tree")


if __name__ == "__main__":
    main()
