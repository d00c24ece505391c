"""The `bcalm`-compatible command line tool run without a GPU: bcalm_main.cpp linked against the CPU emulation build of the
library (tests/emu).  Re-runs the CLI tests of tests/test_gpu_cli.py (input grammars, output naming and header grammar,
-all-abundance-counts, GFA, -edge-km) and adds the counted-k-mer hand-off (-solid-kmers-in / -solid-kmers-out)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from bcalm_b200 import parity  # noqa: E402
from conftest import load_golden  # noqa: E402
import test_gpu_cli as gpu_cli  # noqa: E402
from test_gpu_cli import run_cli, write_fasta  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the CPU emulation of the library")
CASES = {c["name"]: c for c in load_golden()}


@pytest.fixture(scope="module")
def cli():
    import build_emu
    build_emu.build()
    return build_emu.CLI


# the same test bodies as the GPU suite, collected here without the gpu mark and with the emulated binary
test_cli_matches_reference_golden = gpu_cli.test_cli_matches_reference_golden
test_cli_fastq_gz_list_and_out_prefix = gpu_cli.test_cli_fastq_gz_list_and_out_prefix
test_cli_all_abundance_counts_and_gfa = gpu_cli.test_cli_all_abundance_counts_and_gfa
test_cli_edge_km = gpu_cli.test_cli_edge_km


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.mark.parametrize("name", ["reads_1_k21_a2", "lowcx_11_ACGT_k7_a2", "circular_test1"])
def test_counted_kmers_round_trip(cli, tmp_path, name):
    """reads -> unitigs with -solid-kmers-out, then -solid-kmers-in alone must give the same graph."""
    c = CASES[name]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-solid-kmers-out", "solid.txt")
    lines = (tmp_path / "solid.txt").read_text().split("\n")[:-1]
    assert lines and all(len(l.split()[0]) == c["k"] and int(l.split()[1]) >= c["abundance_min"] for l in lines)
    run_cli(cli, tmp_path, "-solid-kmers-in", "solid.txt", "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "again")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "again.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_counted_kmers_from_a_non_canonical_dump(cli, tmp_path):
    """A counter run without canonical counting lists both orientations: counts are added, the abundance filter comes after;
    tabs, CRLF, comments and lower case are tolerated."""
    c = CASES["reads_1_k21_a2"]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", "1", "-solid-kmers-out", "all.txt")
    out = ["# kmer\tcount"]
    for i, l in enumerate((tmp_path / "all.txt").read_text().split("\n")[:-1]):
        km, n = l.split()[0], int(l.split()[1])
        if n >= 2 and i % 2 == 0:
            out.append(f"{km.lower()}\t{n - 1}\r")       # split over the two orientations
            out.append(f"{_rc(km)}\t1")
        else:
            out.append(f"{_rc(km) if i % 3 == 0 else km} {n}")
    (tmp_path / "dump.txt").write_text("\n".join(out) + "\n")
    run_cli(cli, tmp_path, "-solid-kmers-in", "dump.txt", "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-out", "fromdump")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromdump.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_counted_kmers_errors_are_reported_like_the_reference(cli, tmp_path):
    (tmp_path / "bad.txt").write_text("ACGTACGTAC 3\n")
    p = subprocess.run([cli, "-solid-kmers-in", "bad.txt", "-kmer-size", "21"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode != 0 and "EXCEPTION:" in p.stdout and "k-mer length" in p.stdout
    p = subprocess.run([cli, "-solid-kmers-in", "bad.txt", "-kmer-size", "63"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode != 0 and "EXCEPTION:" in p.stdout


@pytest.mark.parametrize("gpus", [2, 3, 4])
@pytest.mark.parametrize("name", ["reads_1_k21_a2", "circular_test1", "even_k_dense_k6", "lowcx_11_ACGT_k7_a2"])
def test_cli_multi_gpu_mode(cli, tmp_path, name, gpus):
    """-nb-gpus N: one context per (emulated) GPU driven by threads of the CLI process, NCCL stand-in between them; the output
    file must be the same graph with consecutive ids, whatever the number of shards."""
    c = CASES[name]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"], width=50)
    out = run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]),
                  "-nb-gpus", str(gpus), "-out-gfa", "g.gfa")
    assert f"GPUs: {gpus}" in out
    text = (tmp_path / "in.unitigs.fa").read_text()
    ids = [int(line[1:].split()[0]) for line in text.split("\n") if line.startswith(">")]
    assert ids == list(range(len(ids)))
    d = parity.diff(parity.parse_unitigs_fa(text), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d
    gfa = (tmp_path / "g.gfa").read_text().split("\n")
    assert sum(l.startswith("S\t") for l in gfa) == len(ids)


def test_cli_multi_gpu_mode_refuses_what_the_sharded_path_cannot_do(cli, tmp_path):
    (tmp_path / "x.fa").write_text(">r\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
    for extra in (["-kmer-size", "41"], ["-kmer-size", "21", "-all-abundance-counts"]):
        p = subprocess.run([cli, "-in", "x.fa", "-nb-gpus", "2", *extra], cwd=tmp_path, capture_output=True, text=True)
        assert p.returncode != 0 and "EXCEPTION:" in p.stdout
