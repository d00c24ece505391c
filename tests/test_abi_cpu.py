"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/cdbg.h declares, the
ctypes structs mirror the C layouts, errors are reported without a GPU, and the product never imports the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cdbg.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cdbg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["cdbg_create", "cdbg_destroy", "cdbg_push_reads", "cdbg_count", "cdbg_compact", "cdbg_link", "cdbg_get_unitigs",
              "cdbg_build", "cdbg_last_error"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from bcalm_b200 import build
    lib = build.build_lib()
    L = C.CDLL(lib)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/cdbg.h but not exported by {lib}"
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cdbg_[a-z_0-9]+)", out))
    assert set(declared_symbols()) <= exported


def test_struct_layouts_match_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from bcalm_b200 import api
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "cdbg.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(cdbg_params), sizeof(cdbg_stats), sizeof(cdbg_link_t), sizeof(cdbg_unitig_view));
  printf("%zu %zu %zu\n", offsetof(cdbg_stats, n_count_bins), offsetof(cdbg_stats, ms_push), offsetof(cdbg_unitig_view, link_off));
  return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c])
        a, b = subprocess.check_output([exe], text=True).strip().split("\n")
    sizes = list(map(int, a.split()))
    offs = list(map(int, b.split()))
    assert sizes == [C.sizeof(api.Params), C.sizeof(api.Stats), C.sizeof(api._Link), C.sizeof(api.UnitigView)]
    assert offs == [api.Stats.n_count_bins.offset, api.Stats.ms_push.offset, api.UnitigView.link_off.offset]
    assert api.LINK_DTYPE.itemsize == C.sizeof(api._Link)


def test_create_fails_loudly_without_gpu_or_with_bad_k():
    import torch
    from bcalm_b200.api import Cdbg, CdbgError
    with pytest.raises(CdbgError) as e:
        Cdbg(k=128)         # reference span list 32/64/96/128: k < 128 ("unhandled kmer size 128")
    assert "unhandled kmer size" in str(e.value)
    with pytest.raises(CdbgError):
        Cdbg(k=3)
    if not torch.cuda.is_available():
        with pytest.raises(CdbgError) as e:
            Cdbg(k=31)
        assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under bcalm_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bcalm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in text and "import oracle" not in text and "liboracle" not in text, f
                assert "#include \"../../oracle" not in text


def test_cli_version_and_errors():
    from bcalm_b200 import build
    build.build_lib()
    cli = build.build_cli()
    out = subprocess.run([cli, "-version"], capture_output=True, text=True)
    assert out.returncode == 0 and "BCALM 2" in out.stdout
    out = subprocess.run([cli], capture_output=True, text=True)
    assert out.returncode != 0 and "EXCEPTION" in out.stdout
    out = subprocess.run([cli, "-in", "/nonexistent.fa", "-bogus-option", "1"], capture_output=True, text=True)
    assert out.returncode != 0 and "EXCEPTION" in out.stdout
