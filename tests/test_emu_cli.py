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


@pytest.mark.parametrize("name", ["reads_1_k21_a2", "lowcx_11_ACGT_k7_a2", "circular_test1", "wide_0_k32_a2", "wide_2_k47_a2", "wide_4_k64_a2", "wide_5_k65_a2",
                                  "wide_7_k96_a2", "wide_8_k127_a2", "wide_lowcx_1_AT_k35"])
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
    assert p.returncode != 0 and "EXCEPTION:" in p.stdout and "k-mer length" in p.stdout


@pytest.mark.parametrize("gpus", [2, 3, 4])
@pytest.mark.parametrize("name", ["reads_1_k21_a2", "circular_test1", "even_k_dense_k6", "lowcx_11_ACGT_k7_a2", "wide_2_k47_a2", "wide_9_k127_a1"])
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


@pytest.mark.parametrize("gpus", [2, 3])
def test_cli_multi_gpu_all_abundance_counts(cli, tmp_path, gpus):
    """-nb-gpus N -all-abundance-counts: the ab:Z: lists against the oracle's per-k-mer abundances, like the single-GPU test."""
    from oracle import oracle as orc
    c = CASES["reads_0_k15_a2"]
    fa = tmp_path / "x.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", "15", "-abundance-min", "2", "-all-abundance-counts", "-nb-gpus", str(gpus))
    recs = parity.parse_unitigs_fa((tmp_path / "x.unitigs.fa").read_text())
    o = parity.graph_to_records(orc.run(c["reads"], 15, 2), with_ab=True)

    def canon(recs_):
        out = []
        for r in recs_:
            s, ab = r["seq"], r["ab"]
            rc = parity.revcomp_str(s)
            assert ab is not None and len(ab) == len(s) - 15 + 1
            out.append((s, tuple(ab)) if s <= rc else (rc, tuple(reversed(ab))))
        return sorted(out)
    assert canon(recs) == canon(o)


def test_cli_multi_gpu_mode_refuses_what_the_sharded_path_cannot_do(cli, tmp_path):
    (tmp_path / "x.fa").write_text(">r\nACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
    for extra in (["-kmer-size", "21", "-kff"], ["-kmer-size", "21", "-gpu-parse", "1"], ["-kmer-size", "21", "-solid-kmers-out", "s.txt"]):
        p = subprocess.run([cli, "-in", "x.fa", "-nb-gpus", "2", *extra], cwd=tmp_path, capture_output=True, text=True)
        assert p.returncode != 0 and "EXCEPTION:" in p.stdout


# ---------------------------------------------------------------------------------------------------------------- KFF
def _kff_read(path, k):
    """Python restatement of the KFF 1.0 layout (kff-cpp-api, vendored by the reference): -> {canonical k-mer string: count}."""
    b = open(path, "rb").read()
    assert b[:3] == b"KFF" and b[-3:] == b"KFF" and b[3] == 1
    enc = b[5]
    nt = {enc >> 6: "A", (enc >> 4) & 3: "C", (enc >> 2) & 3: "G", enc & 3: "T"}
    p = 12 + int.from_bytes(b[8:12], "big")
    var, out = {}, {}
    while p < len(b) - 3:
        t = chr(b[p]); n = int.from_bytes(b[p + 1:p + 9], "big"); p += 9
        if t == "v":
            for _ in range(n):
                z = b.index(0, p)
                var[b[p:z].decode()] = int.from_bytes(b[z + 1:z + 9], "big"); p = z + 9
        elif t == "i":
            p += n * 9 + 8
        elif t == "r":
            assert var["k"] == k
            nbits = max(0, (var["max"] - 1).bit_length())
            nb = (nbits + 7) // 8
            for _ in range(n):
                nk = 1
                if nb:
                    nk = int.from_bytes(b[p:p + nb], "big"); p += nb
                ln = nk + k - 1
                sb = (ln + 3) // 4
                bits = int.from_bytes(b[p:p + sb], "big")
                seq = "".join(nt[(bits >> (2 * (ln - 1 - i))) & 3] for i in range(ln)); p += sb
                for j in range(nk):
                    km = seq[j:j + k]
                    c = int.from_bytes(b[p:p + var["data_size"]], "big") if var["data_size"] else 1
                    p += var["data_size"]
                    out[km] = out.get(km, 0) + c
        else:
            raise AssertionError(f"unexpected section {t!r}")
    return out, var


def _kff_write(path, k, blocks, enc=0x1B, max_kmers=8, data_size=2):
    """A KFF file the way another tool might write it: several k-mers per block (max > 1), another nucleotide encoding, 2-byte counts.
    blocks: list of (sequence, [count per k-mer])."""
    code = {"A": enc >> 6, "C": (enc >> 4) & 3, "G": (enc >> 2) & 3, "T": enc & 3}
    f = bytearray(b"KFF" + bytes([1, 0, enc, 0, 0]) + (0).to_bytes(4, "big"))
    f += b"v" + (3).to_bytes(8, "big")
    for name, v in (("k", k), ("max", max_kmers), ("data_size", data_size)):
        f += name.encode() + b"\0" + v.to_bytes(8, "big")
    f += b"r" + len(blocks).to_bytes(8, "big")
    nb = (max(0, (max_kmers - 1).bit_length()) + 7) // 8
    for seq, counts in blocks:
        nk = len(seq) - k + 1
        assert nk == len(counts) and nk <= max_kmers
        f += nk.to_bytes(nb, "big")
        bits = 0
        for ch in seq:
            bits = (bits << 2) | code[ch]
        f += bits.to_bytes((len(seq) + 3) // 4, "big")
        for c in counts:
            f += c.to_bytes(data_size, "big")
    f += b"KFF"
    open(path, "wb").write(bytes(f))


@pytest.mark.parametrize("name", ["reads_0_k15_a2", "wide_2_k47_a2"])
def test_kff_from_the_reference_is_read(cli, tmp_path, name):
    """tests/golden/<case>.ref.kff were written by the reference binary (`bcalm -kff`, see make_golden.py; one- and two-word k-mers):
    the graph built from each must be the reference's graph of the same reads."""
    c = CASES[name]
    kff = os.path.join(ROOT, "tests", "golden", name + ".ref.kff")
    run_cli(cli, tmp_path, "-solid-kmers-in", kff, "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "fromkff")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromkff.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_kff_written_here_has_the_reference_layout_and_round_trips(cli, tmp_path):
    c = CASES["reads_0_k15_a2"]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-kff", "-solid-kmers-out", "solid.txt")
    ours, var = _kff_read(tmp_path / "in.kff", c["k"])
    ref, ref_var = _kff_read(os.path.join(ROOT, "tests", "golden", "reads_0_k15_a2.ref.kff"), c["k"])
    assert {k_: var[k_] for k_ in ("k", "max", "data_size")} == {k_: ref_var[k_] for k_ in ("k", "max", "data_size")}
    text = dict((l.split()[0], int(l.split()[1])) for l in (tmp_path / "solid.txt").read_text().split("\n")[:-1])
    assert ours == text                                            # the KFF and the text dump hold the same (k-mer, count) pairs
    canon = lambda d_: {min(s, _rc(s)): v for s, v in d_.items()}  # noqa: E731  (the two programs choose the strand differently)
    assert canon(ours) == canon(ref)                               # ... and they are the reference's solid k-mers with its counts
    run_cli(cli, tmp_path, "-solid-kmers-in", "in.kff", "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "again")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "again.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_kff_with_several_kmers_per_block_and_another_encoding(cli, tmp_path):
    c = CASES["reads_1_k21_a2"]
    k = c["k"]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(k), "-abundance-min", str(c["abundance_min"]), "-out", "ref")
    # re-express the unitigs as KFF blocks of up to 8 k-mers with per-k-mer counts from -all-abundance-counts
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(k), "-abundance-min", str(c["abundance_min"]), "-out", "ab", "-all-abundance-counts")
    blocks = []
    lines = (tmp_path / "ab.unitigs.fa").read_text().split("\n")
    for h, s in zip(lines[0::2], lines[1::2]):
        if not h.startswith(">"):
            continue
        ab = [int(x) for x in h.split("ab:Z:")[1].split("L:")[0].split()]
        assert len(ab) == len(s) - k + 1
        for j in range(0, len(ab), 8):
            blocks.append((s[j:j + min(8, len(ab) - j) + k - 1], ab[j:j + 8]))
    _kff_write(tmp_path / "other.kff", k, blocks)
    run_cli(cli, tmp_path, "-solid-kmers-in", "other.kff", "-kmer-size", str(k), "-abundance-min", "1", "-out", "fromother")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromother.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), k)
    assert d == "", d


@pytest.mark.parametrize("name", ["wide_1_k33_a2", "wide_3_k63_a2", "wide_6_k95_a2", "wide_9_k127_a1"])
def test_kff_and_non_canonical_text_dumps_with_multi_word_kmers(cli, tmp_path, name):
    """k = 32..127: the KFF written here holds the same (k-mer, count) pairs as the text dump, both read back into the reference's graph;
    a dump listing the other strand (split counts) is merged like for k <= 31."""
    c = CASES[name]
    k = c["k"]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(k), "-abundance-min", "1", "-kff", "-solid-kmers-out", "all.txt")
    ours, var = _kff_read(tmp_path / "in.kff", k)
    text = dict((l.split()[0], int(l.split()[1])) for l in (tmp_path / "all.txt").read_text().split("\n")[:-1])
    assert var["k"] == k and ours == text and all(len(s_) == k for s_ in text)
    want = parity.parse_unitigs_fa(c["unitigs_fa"])
    run_cli(cli, tmp_path, "-solid-kmers-in", "in.kff", "-kmer-size", str(k), "-abundance-min", str(c["abundance_min"]), "-out", "fromkff")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromkff.unitigs.fa").read_text()), want, k)
    assert d == "", d
    out = []
    for i, (km, n) in enumerate(text.items()):
        if n >= 2 and i % 2 == 0:
            out += [f"{km.lower()}\t{n - 1}", f"{_rc(km)} 1"]
        else:
            out.append(f"{_rc(km) if i % 3 == 0 else km} {n}")
    (tmp_path / "dump.txt").write_text("\n".join(out) + "\n")
    run_cli(cli, tmp_path, "-solid-kmers-in", "dump.txt", "-kmer-size", str(k), "-abundance-min", str(c["abundance_min"]), "-out", "fromdump")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromdump.unitigs.fa").read_text()), want, k)
    assert d == "", d


@pytest.mark.parametrize("chunk,name", [(48, "reads_0_k15_a2"), (900, "reads_1_k21_a2"), (40000, "reads_3_k31_a2")])
def test_large_fasta_is_streamed_to_the_gpu_parser_in_chunks(cli, tmp_path, monkeypatch, chunk, name):
    """Plain FASTA files above a size threshold are mapped, cut at record starts and streamed through two pinned buffers
    (push_fasta_chunked): with tiny chunks -- smaller than a record, about a record, many records -- the graph is the golden one and
    the verbose line says the text was parsed on the GPU."""
    monkeypatch.setenv("CDBG_CLI_CHUNK", str(chunk))
    monkeypatch.setenv("CDBG_CLI_CHUNK_MIN", "0")
    c = CASES[name]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"], width=37)
    out = run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]))
    assert "text parsed on the GPU" in out and f"sequences: {len(c['reads'])} " in out
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "in.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d
    # two files in one run: the second is appended after the first
    fb = tmp_path / "b.fa"
    write_fasta(fb, c["reads"][::-1], width=50)
    run_cli(cli, tmp_path, "-in", f"{fa},{fb}", "-kmer-size", str(c["k"]), "-abundance-min", str(2 * c["abundance_min"]), "-out", "twice")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "twice.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d.count("node multisets differ") <= 1      # same sequences, KC doubled: only the counts may differ
    recs = parity.parse_unitigs_fa((tmp_path / "twice.unitigs.fa").read_text())
    ref = parity.parse_unitigs_fa(c["unitigs_fa"])
    assert sorted(min(r["seq"], parity.revcomp_str(r["seq"])) for r in recs) == sorted(min(r["seq"], parity.revcomp_str(r["seq"])) for r in ref)
    assert sum(r["kc"] for r in recs) == 2 * sum(r["kc"] for r in ref)
