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


def test_product_library_contains_no_emulator_code():
    """The CPU emulation used by the no-GPU tests (tests/emu) is selected by a macro that only its own header defines: the shipped
    library and CLI must contain none of it, and nothing under bcalm_b200/ may include or load anything from tests/."""
    from bcalm_b200 import build
    lib = open(build.build_lib(), "rb").read()
    for marker in (b"emu_ctx_switch", b"[emu]", b"CPU emulation (tests/emu)", b"emu_ncclAllGather"):
        assert marker not in lib, marker
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bcalm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libcdbg_emu" not in text and "_emu_build" not in text, f
                assert "#define CDBG_EMU_ACTIVE" not in text, f


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


def _parse_only(cli, path, *extra):
    out = subprocess.run([cli, "-in", str(path), "-parse-only", *extra], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    f = [l for l in out.stdout.splitlines() if l.startswith("records ")][-1].split()
    return int(f[1]), int(f[3]), f[5]


def test_cli_input_grammar_host_side(tmp_path):
    """FASTA (single/multi-line, CRLF), FASTQ, gz, comma list and file of files all yield the same nucleotide stream
    (G/bank/impl/BankFasta.cpp:390-560; README 'input formats'), checked without a GPU through -parse-only."""
    import gzip
    import numpy as np
    from bcalm_b200 import build
    build.build_lib()
    cli = build.build_cli()
    rng = np.random.default_rng(1)
    reads = ["".join(rng.choice(list("ACGTN"), size=int(rng.integers(30, 300)), p=[.24, .24, .24, .24, .04])) for _ in range(500)]
    expected_bytes = sum(len(r) + 1 for r in reads)
    fa = tmp_path / "a.fa"
    fa.write_text("".join(f">r{i} some ACGT header\n{r}\n" for i, r in enumerate(reads)))
    ml = tmp_path / "ml.fasta"
    ml.write_text("".join(f">r{i}\r\n" + "\r\n".join(r[j:j + 60] for j in range(0, len(r), 60)) + "\r\n" for i, r in enumerate(reads)))
    fq = tmp_path / "a.fq"
    fq.write_text("".join(f"@r{i}\n{r}\n+\n{'@' * len(r)}\n" for i, r in enumerate(reads)))   # quality lines starting with '@'
    gz = tmp_path / "a.fa.gz"
    with gzip.open(gz, "wt") as f:
        f.write(fa.read_text())
    half = tmp_path / "h1.fa"; half2 = tmp_path / "h2.fa"
    half.write_text("".join(f">r{i}\n{r}\n" for i, r in enumerate(reads[:250])))
    half2.write_text("".join(f">r{i}\n{r}\n" for i, r in enumerate(reads[250:])))
    lst = tmp_path / "files.txt"
    lst.write_text("h1.fa\nh2.fa\n")
    ref = _parse_only(cli, fa)
    assert ref[0] == 500 and ref[1] == expected_bytes
    for p in (ml, fq, gz, lst, f"{half},{half2}"):
        assert _parse_only(cli, p) == ref, p


def test_cli_parallel_parse_is_identical(tmp_path):
    import numpy as np
    from bcalm_b200 import build
    cli = build.build_cli()
    rng = np.random.default_rng(2)
    big = tmp_path / "big.fa"
    with open(big, "w") as f:
        for i in range(120000):          # ~12 MB: above the threshold for the multi-threaded parse
            f.write(f">read{i}\n" + "".join(rng.choice(list("ACGT"), size=100)) + "\n")
    a = _parse_only(cli, big, "-nb-cores", "1")
    b = _parse_only(cli, big, "-nb-cores", "7")
    assert a == b and a[0] == 120000
