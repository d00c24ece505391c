#!/usr/bin/env python
"""tests/golden/make_golden.py -- generate golden vectors by running the REAL reference binary.

Run in the build container (needs oracle/_ref/bcalm, built by oracle/build_ref.sh from /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/golden.json: a list of cases {name, k, abundance_min, minimizer_size, reads:[...],
unitigs_fa: "<the reference's unitigs.fa text>"}.  Inputs are either the reference's own fixtures
(R/example/tiny_read.fa, R/test/minitip.fa, R/example/circular_unitigs_unittests/test{1,2,3}.fa,
R/example/pufferize/refs.fa -- the sequences are read from /root/reference at generation time and stored verbatim as
test INPUT data, like any other FASTA handed to the tool) or seeded synthetic inputs generated here.

The committed JSON travels to the GPU box; /root/reference and the script's scratch files do not.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

R = os.environ.get("CDBG_REFERENCE_DIR", "/root/reference")


def read_fasta_seqs(path):
    seqs, cur = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur))
                cur = []
            elif line:
                cur.append(line)
    if cur:
        seqs.append("".join(cur))
    return seqs


def rc(s):
    return s.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]


def synth_reads(rng, genome_len, read_len, coverage, err=0.01, alphabet="ACGT", n_rate=0.0, lower_rate=0.0, circular=False):
    g = "".join(rng.choice(list(alphabet), size=genome_len))
    src = g + g[: read_len - 1] if circular else g
    n_reads = max(1, int(genome_len * coverage / read_len))
    reads = []
    for _ in range(n_reads):
        s = int(rng.integers(0, len(src) - read_len + 1))
        r = list(src[s:s + read_len])
        for i in range(read_len):
            if rng.random() < err:
                r[i] = rng.choice([c for c in "ACGT" if c != r[i]])
            if n_rate and rng.random() < n_rate:
                r[i] = "N"
            if lower_rate and rng.random() < lower_rate:
                r[i] = r[i].lower()
        r = "".join(r)
        if rng.random() < 0.5:
            r = rc(r)
        reads.append(r)
    return reads


def cases():
    out = []
    fx = [
        ("tiny_read_k21", f"{R}/example/tiny_read.fa", 21, 1, 10),
        ("tiny_read_k13", f"{R}/example/tiny_read.fa", 13, 1, 10),
        ("minitip_k21_a1", f"{R}/test/minitip.fa", 21, 1, 10),
        ("minitip_k21_a2", f"{R}/test/minitip.fa", 21, 2, 10),
        ("circular_test1", f"{R}/example/circular_unitigs_unittests/test1.fa", 7, 1, 5),
        ("circular_test2_polyA", f"{R}/example/circular_unitigs_unittests/test2.fa", 7, 1, 5),
        ("circular_test3_onebucket", f"{R}/example/circular_unitigs_unittests/test3.fa", 7, 1, 5),
        ("pufferize_refs", f"{R}/example/pufferize/refs.fa", 9, 1, 5),
    ]
    for name, path, k, a, m in fx:
        out.append(dict(name=name, k=k, abundance_min=a, minimizer_size=m, reads=read_fasta_seqs(path), source=path.replace(R, "R")))
    rng = np.random.default_rng(20261017)
    # low-complexity genomes: dense palindromes / hairpins / self-loops / cycles
    for i, (alpha, k, glen, a) in enumerate([
        ("AT", 5, 40, 1), ("AT", 5, 200, 1), ("AT", 7, 300, 1), ("AT", 9, 600, 1), ("AT", 7, 150, 2),
        ("ACG", 5, 120, 1), ("ACGT", 5, 300, 1), ("ACGT", 6, 400, 1), ("ACGT", 8, 2000, 1), ("AC", 9, 500, 1),
        ("AT", 11, 3000, 1), ("ACGT", 7, 3000, 2), ("AT", 6, 100, 1), ("AT", 8, 400, 1), ("ACGT", 4, 60, 1),
    ]):
        reads = synth_reads(rng, glen, min(glen, 50), 8 if a > 1 else 3, err=0.0 if a == 1 else 0.02, alphabet=alpha)
        out.append(dict(name=f"lowcx_{i}_{alpha}_k{k}_a{a}", k=k, abundance_min=a, minimizer_size=min(5, k - 1), reads=reads))
    # circular genomes (plasmid-like): whole-genome cycles
    for i, (k, glen) in enumerate([(7, 60), (11, 300), (21, 1000), (31, 2500)]):
        reads = synth_reads(rng, glen, min(glen, 80), 12, err=0.0, circular=True)
        out.append(dict(name=f"circular_genome_{i}_k{k}", k=k, abundance_min=1, minimizer_size=min(10, k - 1), reads=reads))
    # sequencing-like
    for i, (k, glen, rl, cov, a, nr, lr) in enumerate([
        (15, 3000, 60, 20, 2, 0.0, 0.0), (21, 5000, 80, 25, 2, 0.002, 0.0), (31, 8000, 100, 30, 2, 0.0, 0.01),
        (31, 20000, 100, 30, 2, 0.001, 0.0), (31, 20000, 150, 30, 3, 0.0, 0.0), (25, 10000, 100, 15, 1, 0.0, 0.0),
        (31, 40, 40, 3, 1, 0.0, 0.0), (31, 31, 31, 2, 1, 0.0, 0.0),
    ]):
        reads = synth_reads(rng, glen, rl, cov, err=0.01, n_rate=nr, lower_rate=lr)
        out.append(dict(name=f"reads_{i}_k{k}_a{a}", k=k, abundance_min=a, minimizer_size=min(10, k - 1), reads=reads))
    # degenerate inputs
    out.append(dict(name="short_reads_only", k=21, abundance_min=1, minimizer_size=10,
                    reads=["ACGTACGTAC", "ACGGTTACGATCGATCGATTTACG", "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"]))
    out.append(dict(name="palindrome_k_even", k=6, abundance_min=1, minimizer_size=5,
                    reads=["ACGCGTAAACGCGTTT", "TTAATTAATTAA", "GAATTCGAATTC"]))
    # cycles whose length is a power of two (pointer doubling lands an element on itself: regression for the K6 tail marker)
    rng2 = np.random.default_rng(77)
    for L, k in [(2, 5), (4, 5), (8, 7), (16, 9), (32, 11), (64, 13), (128, 15), (256, 21), (1024, 31)]:
        if L == 2:
            g = "AC"            # ACACA / CACAC: a 2-cycle of k-mers
        else:
            g = "".join(rng2.choice(list("ACGT"), size=L))
        reads = [(g * (2 + (k + L) // L))[: L + k - 1 + 3], rc((g * (2 + (k + L) // L))[1: L + k + 5])]
        out.append(dict(name=f"pow2_cycle_L{L}_k{k}", k=k, abundance_min=1, minimizer_size=min(10, k - 1), reads=reads))
    # multi-word k-mers (reference spans 64 / 96 / 128): k = 32 ... 127
    rng3 = np.random.default_rng(99)
    for i, (k, glen, rl, cov, a, err) in enumerate([
        (32, 4000, 100, 20, 2, 0.01), (33, 4000, 100, 20, 2, 0.01), (47, 6000, 120, 25, 2, 0.01), (63, 6000, 150, 25, 2, 0.01),
        (64, 6000, 150, 25, 2, 0.01), (65, 6000, 150, 25, 2, 0.01), (95, 8000, 200, 30, 2, 0.005), (96, 8000, 200, 30, 2, 0.005),
        (127, 10000, 250, 30, 2, 0.005), (127, 3000, 250, 12, 1, 0.0), (41, 300, 120, 10, 1, 0.0),
    ]):
        reads = synth_reads(rng3, glen, rl, cov, err=err, n_rate=0.001 if i % 3 == 0 else 0.0, circular=(i == 10))
        out.append(dict(name=f"wide_{i}_k{k}_a{a}", k=k, abundance_min=a, minimizer_size=10, reads=reads))
    # low-complexity with a wide k: palindromic (k-1)-mers need an odd k
    for i, (alpha, k, glen) in enumerate([("AT", 33, 400), ("AT", 35, 600), ("ACG", 37, 500)]):
        reads = synth_reads(rng3, glen, min(glen, 120), 4, err=0.0, alphabet=alpha)
        out.append(dict(name=f"wide_lowcx_{i}_{alpha}_k{k}", k=k, abundance_min=1, minimizer_size=10, reads=reads))
    # even k, dense graphs: a k-mer that is its own reverse complement has BOTH ends on the same side of one junction, so a
    # junction side can carry 5 ends (4 extensions, the self-complementary one twice).  Found by fuzzing the CUDA path against
    # the oracle: the join tables used to keep 4 ids per side.  (own generator: the cases above keep their random streams)
    rng4 = np.random.default_rng(2024)
    for name, alpha, k, glen, m in [("even_k_dense_k4", "ACGT", 4, 3000, 3), ("even_k_dense_k6", "ACGT", 6, 30000, 5),
                                    ("even_k_dense_AT_k8", "AT", 8, 3000, 5), ("even_k_dense_ACG_k6", "ACG", 6, 5000, 5)]:
        g = "".join(rng4.choice(list(alpha), size=glen))
        out.append(dict(name=name, k=k, abundance_min=1, minimizer_size=m, reads=[g[: glen // 2 + k], rc(g[glen // 2:])]))
    reads = synth_reads(rng4, 6000, 90, 25, err=0.01)
    out.append(dict(name="even_k_reads_k20_a2", k=20, abundance_min=2, minimizer_size=10, reads=reads))
    # the literal inputs of the reference's own GraphUnitigs unit tests (TestDebruijnUnitigs.cpp:404-735), with their flags
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_reference_unit_vectors import VECTORS
    for name, reads, k, _ in VECTORS:
        out.append(dict(name=f"ut_{name}_k{k}", k=k, abundance_min=1, minimizer_size=3 if k == 5 or name == "test12" else 10, reads=list(reads)))
    out.append(dict(name="ut_test12_normal_min_k21", k=21, abundance_min=1, minimizer_size=10, reads=list(VECTORS[5][1])))
    out.append(dict(name="ut_test4_k27", k=27, abundance_min=1, minimizer_size=10, reads=["ACCATGTATAATTATAAGTAGGTACCT"]))
    return out


def main():
    if not orc.have_ref():
        sys.exit("oracle/_ref/bcalm missing: run oracle/build_ref.sh first")
    res = []
    for c in cases():
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "in.fa")
            orc.write_fasta(fa, c["reads"])
            try:
                text, dt, log = orc.run_ref(fa, c["k"], c["abundance_min"], nb_cores=2,
                                            extra=["-minimizer-size", str(c["minimizer_size"]), "-verbose", "0"], timeout=300)
            except Exception as e:  # the reference refuses some degenerate inputs; record that too
                text = None
                c["ref_error"] = str(e)[:300]
        c["unitigs_fa"] = text
        n = text.count(">") if text else -1
        print(f"{c['name']:32s} k={c['k']:2d} a={c['abundance_min']} reads={len(c['reads']):5d} -> {n} unitigs")
        res.append(c)
    # the reference's own KFF dumps (-kff) of two cases, one- and two-word k-mers: fixtures for the KFF reader of the CLI (tests/test_emu_cli.py)
    write_reference_kff(res)
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(res, f, indent=0)
    print("wrote tests/golden/golden.json", os.path.getsize(os.path.join(ROOT, "tests", "golden", "golden.json")), "bytes")


def write_reference_kff(cases, names=("reads_0_k15_a2", "wide_2_k47_a2")):
    for name in names:
        kc = [c for c in cases if c["name"] == name][0]
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "r.fa")
            orc.write_fasta(fa, kc["reads"])
            extra = ["-minimizer-size", str(kc["minimizer_size"])] if kc.get("minimizer_size") else []
            orc.run_ref(fa, kc["k"], kc["abundance_min"], nb_cores=2, extra=[*extra, "-verbose", "0", "-kff"], timeout=300, keep_dir=d)
            for fn in os.listdir(d):
                if fn.endswith(".kff"):
                    with open(os.path.join(d, fn), "rb") as fi, open(os.path.join(ROOT, "tests", "golden", name + ".ref.kff"), "wb") as fo:
                        fo.write(fi.read())


if __name__ == "__main__":
    main()
