"""GPU tests of the `bcalm`-compatible CLI: input grammars (FASTA multi-line, FASTQ, gz, comma list, file of files), output
file naming and header grammar, compared with the reference's golden unitigs.fa under the parity relation."""
import gzip
import os
import subprocess

import pytest

from bcalm_b200 import build, parity
from conftest import load_golden

pytestmark = pytest.mark.gpu
CASES = {c["name"]: c for c in load_golden()}


@pytest.fixture(scope="module")
def cli():
    build.build_lib()
    return build.build_cli()


def run_cli(cli, cwd, *args):
    p = subprocess.run([cli, *args], cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def write_fasta(path, reads, width=None):
    with open(path, "w") as f:
        for i, r in enumerate(reads):
            f.write(f">read{i} ACGT comment\n")
            if width:
                for j in range(0, len(r), width):
                    f.write(r[j:j + width] + "\n")
            else:
                f.write(r + "\n")


@pytest.mark.parametrize("name", ["tiny_read_k21", "minitip_k21_a1", "circular_test1", "circular_test2_polyA", "pufferize_refs",
                                  "lowcx_11_ACGT_k7_a2", "reads_3_k31_a2", "reads_1_k21_a2", "wide_3_k63_a2", "wide_8_k127_a2"])
def test_cli_matches_reference_golden(cli, tmp_path, name):
    c = CASES[name]
    fa = tmp_path / "input.reads.fa"
    write_fasta(fa, c["reads"], width=37)          # multi-line FASTA
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]),
            "-minimizer-size", str(c["minimizer_size"]), "-max-memory", "1000", "-nb-cores", "2")
    out = tmp_path / "input.reads.unitigs.fa"      # basename cut at the LAST dot, in the CWD
    assert out.exists()
    d = parity.diff(parity.parse_unitigs_fa(out.read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_cli_fastq_gz_list_and_out_prefix(cli, tmp_path):
    c = CASES["reads_2_k31_a2"]
    reads = c["reads"]
    half = len(reads) // 2
    fq = tmp_path / "a.fq"
    with open(fq, "w") as f:
        for i, r in enumerate(reads[:half]):
            f.write(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n")
    gz = tmp_path / "b.fa.gz"
    with gzip.open(gz, "wt") as f:
        for i, r in enumerate(reads[half:]):
            f.write(f">r{i}\n{r}\n")
    # comma separated list
    run_cli(cli, tmp_path, "-in", f"{fq},{gz}", "-kmer-size", "31", "-abundance-min", "2", "-out", str(tmp_path / "viaComma"), "-verbose", "0")
    # file of files
    lst = tmp_path / "files.lst"
    lst.write_text("a.fq\nb.fa.gz\n")
    run_cli(cli, tmp_path, "-in", str(lst), "-kmer-size", "31", "-abundance-min", "2", "-out", str(tmp_path / "viaList"))
    ref = parity.parse_unitigs_fa(c["unitigs_fa"])
    for name in ("viaComma.unitigs.fa", "viaList.unitigs.fa"):
        d = parity.diff(parity.parse_unitigs_fa((tmp_path / name).read_text()), ref, 31)
        assert d == "", d


def test_cli_all_abundance_counts_and_gfa(cli, tmp_path):
    from oracle import oracle as orc
    c = CASES["reads_0_k15_a2"]
    fa = tmp_path / "x.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", "15", "-abundance-min", "2", "-all-abundance-counts", "-out-gfa", str(tmp_path / "x.gfa"))
    recs = parity.parse_unitigs_fa((tmp_path / "x.unitigs.fa").read_text())
    o = parity.graph_to_records(orc.run(c["reads"], 15, 2), with_ab=True)
    # ab:Z: lists: compare as multisets of (canonical sequence, abundance list in canonical orientation)
    def canon(recs_):
        out = []
        for r in recs_:
            s, ab = r["seq"], r["ab"]
            rc = parity.revcomp_str(s)
            assert ab is not None and len(ab) == len(s) - 15 + 1
            out.append((s, tuple(ab)) if s <= rc else (rc, tuple(reversed(ab))))
        return sorted(out)
    assert canon(recs) == canon(o)
    gfa = (tmp_path / "x.gfa").read_text().splitlines()
    assert gfa[0] == "H\tVN:Z:1.0\tks:i:15"
    assert sum(1 for l in gfa if l.startswith("S\t")) == len(recs)
    assert sum(1 for l in gfa if l.startswith("L\t")) == sum(len(r["links"]) for r in recs)
    assert all(l.endswith("\t14M") for l in gfa if l.startswith("L\t"))


def test_cli_edge_km(cli, tmp_path):
    """-edge-km 1 prints J:<0|1>:<id>:<0|1> instead of L: (LinkTigs.cpp:380-384,440-444): same links, other spelling."""
    c = CASES["lowcx_11_ACGT_k7_a2"]
    fa = tmp_path / "e.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", "7", "-abundance-min", "2", "-edge-km", "1", "-out", str(tmp_path / "j"))
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", "7", "-abundance-min", "2", "-out", str(tmp_path / "l"))
    import re
    jt = (tmp_path / "j.unitigs.fa").read_text()
    assert " L:" not in jt
    conv = {("0", "1"): "L:-:%s:-", ("0", "0"): "L:-:%s:+", ("1", "0"): "L:+:%s:+", ("1", "1"): "L:+:%s:-"}
    jt2 = re.sub(r"J:([01]):(\d+):([01])", lambda m: conv[(m.group(1), m.group(3))] % m.group(2), jt)
    d = parity.diff(parity.parse_unitigs_fa(jt2), parity.parse_unitigs_fa((tmp_path / "l.unitigs.fa").read_text()), 7)
    assert d == "", d


# ---------------------------------------------------------------------------------------------------------------- counted k-mers in / out, KFF, GFA
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.mark.parametrize("name", ["reads_1_k21_a2", "lowcx_11_ACGT_k7_a2", "circular_test1", "wide_2_k47_a2", "wide_8_k127_a2"])
def test_cli_counted_kmers_round_trip(cli, tmp_path, name):
    """reads -> unitigs with -solid-kmers-out, then -solid-kmers-in alone (stages 2-4 only) must give the reference's graph
    (GraphUnitigs.cpp:884-951: a counted input skips the counting stage)."""
    c = CASES[name]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-solid-kmers-out", "solid.txt")
    lines = (tmp_path / "solid.txt").read_text().split("\n")[:-1]
    assert lines and all(len(l.split()[0]) == c["k"] and int(l.split()[1]) >= c["abundance_min"] for l in lines)
    run_cli(cli, tmp_path, "-solid-kmers-in", "solid.txt", "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "again")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "again.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


@pytest.mark.parametrize("name", ["reads_0_k15_a2", "wide_2_k47_a2"])
def test_cli_kff_written_by_the_reference_is_read(cli, tmp_path, name):
    """tests/golden/<case>.ref.kff were written by the reference binary (`bcalm -kff`; one- and two-word k-mers): the graph built from
    each must be the reference's graph of the same reads."""
    c = CASES[name]
    kff = os.path.join(ROOT, "tests", "golden", name + ".ref.kff")
    run_cli(cli, tmp_path, "-solid-kmers-in", kff, "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "fromkff")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "fromkff.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_cli_kff_round_trip(cli, tmp_path):
    c = CASES["reads_0_k15_a2"]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-kff")
    assert (tmp_path / "in.kff").read_bytes()[:3] == b"KFF"
    run_cli(cli, tmp_path, "-solid-kmers-in", "in.kff", "-kmer-size", str(c["k"]), "-abundance-min", "1", "-out", "again")
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "again.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


def test_cli_gfa_is_what_convert_to_gfa_makes_of_the_fasta(cli, tmp_path):
    """-out-gfa against the reference's own converter applied to OUR unitigs.fa: scripts/convertToGFA.py:74,97-114 writes
    H VN:Z:1.0 ks:i:k, one S line per record (id, sequence, the LN/KC/km tags) and one L line per L: token (from, sign, to, sign,
    (k-1)M).  The expected line set is derived here from the FASTA the same way."""
    c = CASES["reads_3_k31_a2"]
    k = c["k"]
    fa = tmp_path / "x.fa"
    write_fasta(fa, c["reads"])
    run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(k), "-abundance-min", str(c["abundance_min"]), "-out-gfa", str(tmp_path / "x.gfa"))
    expect = {f"H\tVN:Z:1.0\tks:i:{k}"}
    text = (tmp_path / "x.unitigs.fa").read_text().split("\n")
    for h, s in zip(text[0::2], text[1::2]):
        if not h.startswith(">"):
            continue
        toks = h[1:].split()
        tags = [t for t in toks[1:] if t[:2] in ("LN", "KC", "km")]
        expect.add("\t".join(["S", toks[0], s] + tags))
        for t in toks[1:]:
            if t.startswith("L:"):
                _, s1, to, s2 = t.split(":")
                expect.add("\t".join(["L", toks[0], s1, to, s2, f"{k - 1}M"]))
    got = (tmp_path / "x.gfa").read_text().split("\n")
    assert got[-1] == "" and len(got) - 1 == len(set(got[:-1]))
    assert set(got[:-1]) == expect


def test_cli_gpu_parser_and_host_parser_give_the_same_file(cli, tmp_path):
    """The default parses FASTA / FASTQ text on the GPU (cdbg_push_sequence_text); -gpu-parse 0 is the host parser
    (BankFasta.cpp:390-560 semantics): multi-line FASTA with lower case and N's, and FASTQ."""
    c = CASES["reads_3_k31_a2"]
    fa = tmp_path / "m.fa"
    write_fasta(fa, [r if i % 5 else r[:40].lower() + "N" + r[41:] for i, r in enumerate(c["reads"])], width=41)
    fq = tmp_path / "m.fq"
    with open(fq, "w") as f:
        for i, r in enumerate(c["reads"]):
            f.write(f"@r{i} x\n{r}\n+\n{'F' * len(r)}\n")
    for src in (fa, fq):
        outs = []
        for mode in ("1", "0"):
            run_cli(cli, tmp_path, "-in", str(src), "-kmer-size", "31", "-abundance-min", "2", "-gpu-parse", mode, "-out", f"p{mode}")
            outs.append(parity.parse_unitigs_fa((tmp_path / f"p{mode}.unitigs.fa").read_text()))
        assert parity.diff(outs[0], outs[1], 31) == ""
    d = parity.diff(parity.parse_unitigs_fa((tmp_path / "p1.unitigs.fa").read_text()), parity.parse_unitigs_fa(c["unitigs_fa"]), 31)
    assert d == "", d                    # (the FASTQ holds the unmodified reads)


@pytest.mark.parametrize("name", ["reads_1_k21_a2", "circular_test1", "wide_2_k47_a2"])
def test_cli_multi_gpu_mode(cli, tmp_path, name):
    """-nb-gpus 2: one context per GPU driven by threads of the CLI process, NCCL between them; one output file, consecutive ids."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    c = CASES[name]
    fa = tmp_path / "in.fa"
    write_fasta(fa, c["reads"], width=50)
    out = run_cli(cli, tmp_path, "-in", str(fa), "-kmer-size", str(c["k"]), "-abundance-min", str(c["abundance_min"]), "-nb-gpus", "2")
    assert "GPUs: 2" in out
    text = (tmp_path / "in.unitigs.fa").read_text()
    ids = [int(line[1:].split()[0]) for line in text.split("\n") if line.startswith(">")]
    assert ids == list(range(len(ids)))
    d = parity.diff(parity.parse_unitigs_fa(text), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d
