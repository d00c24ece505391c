"""GPU-side FASTA / FASTQ parsing (parse.cuh, cdbg_push_sequence_text) checked under the CPU emulator: the nucleotide stream
it produces must be byte-identical to what the host parser of the CLI produces (restated here in Python from
parse_sequences_raw in bcalm_main.cpp), for multi-line FASTA, CRLF, headers full of ACGT, blank lines, records longer than a
tile, FASTQ with quality strings that look like headers; text it cannot parse exactly must be refused."""
from __future__ import annotations

import os
import random
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from bcalm_b200 import api, parity  # noqa: E402
from conftest import load_golden  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the CPU emulation of the library")
CASES = {c["name"]: c for c in load_golden()}


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield build_emu
    api.LIB_PATH, api._lib = saved


def host_parse(text: str):
    """Python restatement of parse_sequences_raw (bcalm_main.cpp): -> (nucleotide stream, number of records)."""
    lines = text.split("\n")
    if text.endswith("\n"):
        lines = lines[:-1]
    out, n, i = [], 0, 0
    strip = lambda s: s[:-1] if s.endswith("\r") else s  # noqa: E731
    while i < len(lines):
        if lines[i].startswith(">"):
            i += 1
            seq = ""
            while i < len(lines) and not lines[i].startswith(">"):
                seq += strip(lines[i])
                i += 1
            out.append(seq + "\n")
            n += 1
        elif lines[i].startswith("@"):
            out.append(strip(lines[i + 1] if i + 1 < len(lines) else "") + "\n")
            n += 1
            i += 4
        else:
            i += 1
    return "".join(out), n


def _texts():
    rng = random.Random(3)
    rnd = lambda n, alphabet="ACGTN": "".join(rng.choice(alphabet) for _ in range(n))  # noqa: E731
    fa = ""
    for r in range(400):
        s, w = rnd(rng.randint(0, 300)), rng.choice([7, 60, 80, 1000])
        fa += f">read{r} ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT len={len(s)}" + rng.choice(["\n", "\r\n"])
        for j in range(0, len(s), w):
            fa += s[j:j + w] + rng.choice(["\n", "\r\n"])
        if rng.random() < 0.1:
            fa += "\n"
    fq = ""
    for r in range(500):
        s = rnd(rng.randint(1, 250), "ACGTacgtN")
        fq += f"@r{r} x\n{s}\n+\n{rnd(len(s), 'ACGT@+>IF#')}\n"
    return {
        "fasta_multiline_crlf_blank_lines": fa,
        "fasta_no_trailing_newline": fa.rstrip("\r\n"),
        "fasta_ends_with_a_header": fa + ">lonely header",
        "fasta_one_record_over_many_tiles": ">g\n" + ("ACGTTGCA" * 20 + "\n") * 3000,
        "fasta_single_line_genome": ">chr\n" + rnd(100_000, "ACGT") + "\n",
        "fasta_empty_records": ">a\n>b\n>c\nACGT\n>d\n",
        "fastq": fq,
        "fastq_crlf": fq.replace("\n", "\r\n"),
        "fastq_no_trailing_newline": fq[:-1],
    }


TEXTS = _texts()


@pytest.mark.parametrize("name", sorted(TEXTS))
def test_gpu_parse_equals_host_parse(emu_lib, name):
    text = TEXTS[name]
    with api.Cdbg(k=21, abundance_min=1, device=0) as c:
        n = c.push_sequence_text(text)
        got = c.pushed_reads().decode()
    want, n_want = host_parse(text)
    assert n == n_want
    assert got == want


def test_text_that_cannot_be_parsed_exactly_is_refused(emu_lib):
    multi_line_fastq = "@r1\nACGT\nACGT\n+\nIIII\nIIII\n@r2\nACGT\n+\nIIII\n"
    truncated = "@r1\nACGT\n+\nIIII\n@r2\nACG"
    with api.Cdbg(k=21, abundance_min=1, device=0) as c:
        for bad in (multi_line_fastq, truncated, "no header here\nACGT\n"):
            with pytest.raises(api.CdbgError):
                c.push_sequence_text(bad)
        assert c.pushed_reads() == b""                       # nothing was appended
        assert c.push_sequence_text(">ok\nACGT\n") == 1      # and the context is still usable


def test_several_pushes_and_mixed_entry_points(emu_lib):
    c = CASES["reads_1_k21_a2"]
    reads = c["reads"]
    third = len(reads) // 3
    fa = "".join(f">r{i}\n{r[:40]}\n{r[40:]}\n" for i, r in enumerate(reads[:third]))
    fq = "".join(f"@q{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads[third:2 * third]))
    with api.Cdbg(k=c["k"], abundance_min=c["abundance_min"], device=0) as x:
        assert x.push_sequence_text(fa) == third
        assert x.push_sequence_text(fq) == third
        x.push_reads(reads[2 * third:])
        x.count(); x.compact(); x.link()
        g = x.unitigs()
    d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


@pytest.mark.parametrize("name", ["reads_1_k21_a2", "minitip_k21_a1", "wide_3_k63_a2"])
def test_cli_with_gpu_parse_matches_reference_golden(emu_lib, tmp_path, name):
    c = CASES[name]
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, r in enumerate(c["reads"]):
            f.write(f">read{i} ACGTACGTACGTACGTACGTACGTACGTAC\n")
            for j in range(0, len(r), 33):
                f.write(r[j:j + 33] + "\n")
    p = subprocess.run([emu_lib.CLI, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-gpu-parse", "1"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"sequences: {len(c['reads'])} " in p.stdout
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "in.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def _fnv1a(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("seed", [41, 42])
def test_random_texts_against_the_cli_host_parser(emu_lib, tmp_path, seed):
    """Random FASTA / FASTQ formatting (line widths 1..inf, CRLF, blank lines, headers and quality strings full of '>', '@', '+',
    ACGT) against the C++ host parser itself (`bcalm -parse-only` prints record count, byte count and an FNV-1a of its stream)."""
    rng = random.Random(seed)
    for it in range(60):
        fmt = rng.choice(["fa", "fa", "fq"])
        nl = lambda: rng.choice(["\n", "\n", "\n", "\r\n"])  # noqa: E731
        parts = []
        for _ in range(rng.choice([1, 2, 17, 300])):
            length = rng.choice([0, 1, 5, 70, 150, 9000]) if rng.random() < 0.2 else rng.randint(20, 200)
            seq = "".join(rng.choice("ACGTNacgtn") for _ in range(length))
            junk = "".join(rng.choice("ACGT >@+_|:0123") for _ in range(rng.randint(0, 60)))
            if fmt == "fa":
                parts.append(">" + junk + nl())
                w = rng.choice([1, 7, 60, 80, 100000])
                parts += [seq[j:j + w] + nl() for j in range(0, len(seq), w)]
                if rng.random() < 0.05:
                    parts.append(nl())
            else:
                parts += ["@" + junk + nl(), seq + nl(), "+" + nl(), "".join(rng.choice("ACGT@+>!IJ#") for _ in seq) + nl()]
        text = "".join(parts)
        if rng.random() < 0.3:
            text = text.rstrip("\r\n")
        path = tmp_path / f"t{it}.{fmt}"
        with open(path, "w", newline="") as f:
            f.write(text)
        p = subprocess.run([emu_lib.CLI, "-parse-only", "-in", str(path), "-verbose", "0"], capture_output=True, text=True)
        tok = [ln for ln in p.stdout.splitlines() if ln.startswith("records ")][0].split()
        with api.Cdbg(k=21, abundance_min=1, device=0) as c:
            try:
                n = c.push_sequence_text(text)
            except api.CdbgError:
                assert fmt == "fq"          # only FASTQ text may be refused (then the CLI parses it on the host)
                continue
            got = c.pushed_reads()
        assert (n, len(got), _fnv1a(got)) == (int(tok[1]), int(tok[3]), int(tok[5], 16)), (seed, it, fmt)
