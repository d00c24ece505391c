"""Known answers from the reference's own unit tests for this path, restated as k-mer level neighbourhoods.

gatb-core's TestDebruijnUnitigs.cpp builds a GraphUnitigs from a few literal strings and asserts the neighbours of named nodes
(tests 1, 7, 10, 11, 12, 14: G/../test/unit/src/debruijn/TestDebruijnUnitigs.cpp:404-735).  GraphUnitigs is the loader of the very
`unitigs.fa` this path writes (sequences + L: links), so the same neighbourhoods must be derivable from our output: a k-mer's
successor is the next k-mer of its unitig or, at a unitig end, the first k-mer of every linked unitig in the linked orientation.
Checked for the oracle and for the CUDA path under the CPU emulator; the GPU suite checks the same inputs through the golden file
(cases ut_*), whose expected output comes from the reference binary."""
from __future__ import annotations

import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from bcalm_b200 import api, parity  # noqa: E402
from oracle import oracle as orc  # noqa: E402

COMP = str.maketrans("ACGT", "TGCA")


def rc(s):
    return s[::-1].translate(COMP)


def kmer_edges(records, k):
    """All directed k-mer adjacencies x -> y (y = x[1:] + c) spelled by unitigs + links, closed under reverse complement."""
    by_id = {r["id"]: r for r in records}
    edges = set()

    def add(a, b):
        assert a[1:] == b[:-1], (a, b)
        edges.add((a, b)); edges.add((rc(b), rc(a)))
    for r in records:
        s = r["seq"]
        for i in range(len(s) - k):
            add(s[i:i + k], s[i + 1:i + k + 1])
        for s1, v, s2 in r["links"]:
            src = s if s1 == "+" else rc(s)            # leave through the END ('+') or through the BEGIN ('-')
            t = by_id[v]["seq"]
            dst = t if s2 == "+" else rc(t)            # enter at the BEGIN ('+') or at the END ('-')
            add(src[-k:], dst[:k])
    return edges


def out_neighbours(edges, x):
    return {b for a, b in edges if a == x}


def in_neighbours(edges, x):
    return {a for a, b in edges if b == x}


T12 = ["CATCGATGCGAGACGCCTGTCGCGGGGAATTGTGGGGCGGACCACGCTCTGGCTAACGAGCTACCGTTTCCTTTAACCTGCCAGACGGTGACCAGGGCCGTTCGGCGTTGCATCGAGCGGTGTCGCTAGCGCAATGCGCAAGATTTTGACATTTACAAGGCAACATTGCAGCGTCCGATGGTCCGGTGGCCTCCAGATAGTGTCCAGTCGCTCTAACTGTATGGAGACCATAGGCATTTACCTTATTCTCATCGCCACGCCCCAAGATCTTTAGGACCCAGCATTCCTTTAACCACTAACATAACGCGTGTCATCTAGTTCAACAACC",
       "TGTCATCTAGTTCAACAACCAAAAAAA",
       "TGTCATCTAGTTCAACAACCGTTATGCCGTCCGACTCTTGCGCTCGGATGTCCGCAATGGGTTATCCCTATGTTCCGGTAATCTCTCATCTACTAAGCGCCCTAAAGGTCGTATGGTTGGAGGGCGGTTACACACCCTTAAGTACCGAACGATAGAGCACCCGTCTAGGAGGGCGTGCAGGGTCTCCCGCTAGCTAATGGTCACGGCCTCTCTGGGAAAGCTGAACAACGGATGATACCCATACTGCCACTCCAGTACCTGGGCCGCGTGTTGTACGCTGTGTATCTTGAGAGCGTTTCCAGCAGATAGAACAGGATCACATGTACATG"]

# (name, reads, k, [(node, "out"/"in", expected neighbour set)])
VECTORS = [
    ("test1", ["AGGCGA", "TTGCGA", "GCGAT", "GCGAA"], 5,
     [("GGCGA", "out", {"GCGAA", "GCGAT"}), ("GCGAT", "in", {"GGCGA", "TGCGA"})]),
    ("test7", ["AGGCGC", "ACTGACTGACTGACTG"], 5,
     [("AGGCG", "out", {"GGCGC"}), ("GGCGC", "out", {"GCGCC"}), ("GCGCC", "out", {"CGCCT"}), ("ACTGA", "out", {"CTGAC"}), ("GACTG", "out", {"ACTGA"})]),
    ("test7_nocircular", ["AGGCGC", "ACTGACT"], 5, [("AGGCG", "out", {"GGCGC"}), ("GGCGC", "out", {"GCGCC"}), ("GCGCC", "out", {"CGCCT"})]),
    ("test10", ["AGGCGA", "TTGCGA", "GCGAT", "CGATA", "CGATT"], 5,
     [("GCGAT", "out", {"CGATA", "CGATT"}), ("GCGAT", "in", {"GGCGA", "TGCGA"}), ("ATCGC", "out", {"TCGCC", "TCGCA"}), ("ATCGC", "in", {"TATCG", "AATCG"})]),
    ("test11", ["AGGCGA", "TTGCGA", "GCGAT", "CGATA"], 5,
     [("GCGAT", "out", {"CGATA"}), ("GCGAT", "in", {"GGCGA", "TGCGA"}), ("ATCGC", "out", {"TCGCC", "TCGCA"}), ("ATCGC", "in", {"TATCG"})]),
    ("test12", T12, 21,
     [("TGTCATCTAGTTCAACAACCA", "in", {"GTGTCATCTAGTTCAACAACC"}), ("TGGTTGTTGAACTAGATGACA", "out", {"GGTTGTTGAACTAGATGACAC"})]),
    ("test14", ["AAGAGCCCGCTTATCCGGTGGTGATACCTAC", "AGAGCCCGCTTATCCGGTGGTGATACCTACC"], 31,
     [("AGAGCCCGCTTATCCGGTGGTGATACCTACC", "in", {"AAGAGCCCGCTTATCCGGTGGTGATACCTAC"}), ("AGAGCCCGCTTATCCGGTGGTGATACCTACC", "out", set())]),
]


def check(records, k, expectations):
    edges = kmer_edges(records, k)
    for node, direction, want in expectations:
        got = out_neighbours(edges, node) if direction == "out" else in_neighbours(edges, node)
        assert got == want, (node, direction, got, want)


@pytest.mark.parametrize("name,reads,k,expect", VECTORS, ids=[v[0] for v in VECTORS])
def test_oracle_reproduces_the_reference_unit_test_neighbourhoods(name, reads, k, expect):
    check(parity.graph_to_records(orc.run(reads, k, 1)), k, expect)


@pytest.fixture(scope="module")
def emu_lib():
    if shutil.which("g++") is None:
        pytest.skip("needs g++ to build the CPU emulation of the library")
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield path
    api.LIB_PATH, api._lib = saved


@pytest.mark.parametrize("name,reads,k,expect", VECTORS, ids=[v[0] for v in VECTORS])
def test_cuda_path_reproduces_the_reference_unit_test_neighbourhoods(emu_lib, name, reads, k, expect):
    with api.Cdbg(k=k, abundance_min=1, device=0) as c:
        g = c.build(reads)
    check(parity.graph_to_records(g), k, expect)
