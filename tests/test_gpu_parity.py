"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from bcalm_b200 import parity
from bcalm_b200.api import Cdbg
from conftest import load_golden

pytestmark = pytest.mark.gpu

CASES = load_golden()


def _records(g):
    return parity.graph_to_records(g)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_vs_reference_output(case):
    """CUDA path == the real reference binary's unitigs.fa (committed golden), under the parity relation."""
    if case["unitigs_fa"] is None:
        pytest.skip("reference refused this input")
    k = case["k"]
    with Cdbg(k=k, abundance_min=case["abundance_min"], minimizer_size=case["minimizer_size"]) as c:
        g = c.build(case["reads"])
    d = parity.diff(_records(g), parity.parse_unitigs_fa(case["unitigs_fa"]), k)
    assert d == "", d


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_solid_kmers_vs_oracle(case):
    """Stage 1 in isolation: the multiset of solid (k-mer, abundance) pairs is bit-exact."""
    from oracle import oracle as orc
    k = case["k"]
    o = orc.run_any(case["reads"], k, case["abundance_min"], stages=1)
    with Cdbg(k=k, abundance_min=case["abundance_min"]) as c:
        c.push_reads(case["reads"])
        st = c.count()
        km, ct = c.solid_kmers()
    assert st["n_instances"] == o.stats["n_instances"]
    assert st["n_distinct"] == o.stats["n_distinct"]
    if k <= 31:
        order = np.argsort(km, kind="stable")
        assert np.array_equal(km[order], o.solid_kmers)
    else:   # multi-word k-mers: sort lexicographically, most significant word first
        W = km.shape[1]
        order = np.lexsort(tuple(km[:, i] for i in range(W)))
        assert np.array_equal(km[order], o.solid_kmers[:, :W]) and (o.solid_kmers[:, W:] == 0).all()
    assert np.array_equal(ct[order], o.solid_counts)
