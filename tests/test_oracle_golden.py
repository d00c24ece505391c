"""The oracle (oracle/cdbg_oracle.c) must reproduce the REAL reference binary's output on every golden vector
(tests/golden/golden.json, produced by tests/golden/make_golden.py from oracle/_ref/bcalm)."""
import pytest

from bcalm_b200 import parity
from oracle import oracle as orc
from conftest import load_golden

CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    if case["unitigs_fa"] is None:
        pytest.skip("reference refused this input: " + case.get("ref_error", ""))
    k = case["k"]
    g = orc.run_any(case["reads"], k, case["abundance_min"])
    ours = parity.graph_to_records(g)
    ref = parity.parse_unitigs_fa(case["unitigs_fa"])
    d = parity.diff(ours, ref, k)
    assert d == "", d


@pytest.mark.parametrize("case", [c for c in CASES if c["k"] <= 31], ids=[c["name"] for c in CASES if c["k"] <= 31])
def test_wide_oracle_equals_narrow_oracle(case):
    """cdbg_oracle_wide.c (4-word k-mers) must agree with cdbg_oracle.c wherever both apply (k <= 31)."""
    k = case["k"]
    a = orc.run(case["reads"], k, case["abundance_min"])
    b = orc.run_wide(case["reads"], k, case["abundance_min"])
    assert a.stats == b.stats
    assert (b.solid_kmers[:, 1:] == 0).all() and (b.solid_kmers[:, 0] == a.solid_kmers).all() and (a.solid_counts == b.solid_counts).all()
    d = parity.diff(parity.graph_to_records(a), parity.graph_to_records(b), k)
    assert d == "", d


def test_survey_known_answers():
    """Golden strings quoted in SURVEY.md 8c (independent of the JSON)."""
    g = orc.run(["ACTGCTGACTGAGTCATGTGTGGGT"], 21, 1)
    r = parity.graph_to_records(g)
    assert len(r) == 1 and r[0]["ln"] == 25 and r[0]["kc"] == 5 and r[0]["km"] == "1.0"
    assert min(r[0]["seq"], parity.revcomp_str(r[0]["seq"])) == "ACCCACACATGACTCAGTCAGCAGT"
    assert g.stats["n_solid"] == 5


def test_dsk_check1_known_answers():
    """G/../test/unit/src/kmer/TestDSK.cpp:147-175 (DSK_check1): solid k-mer counts for s1 at k=27/26, nks=1..3."""
    s1 = "GATCCTCCCCAGGCCCCTACACCCAAT"
    assert orc.run([s1], 27, 1, stages=1).stats["n_solid"] == 1
    assert orc.run([s1], 26, 1, stages=1).stats["n_solid"] == 2
    assert orc.run([s1, s1], 27, 2, stages=1).stats["n_solid"] == 1
    assert orc.run([s1, s1], 26, 2, stages=1).stats["n_solid"] == 2
    assert orc.run([s1, s1], 27, 3, stages=1).stats["n_solid"] == 0
