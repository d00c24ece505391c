"""Full-size checks at BASELINE.json's bench configuration (C2: 5 Mbp genome, 1.5 M x 100 bp reads, k=31, abundance-min 2).

The C oracle still finishes C2 in well under a minute, so besides the size-independent properties the complete graph is
compared (canonical node and link multisets)."""
import numpy as np
import pytest

from bcalm_b200 import parity
from bcalm_b200.api import Cdbg, MEM_DEVICE

pytestmark = pytest.mark.gpu

K, AMIN = 31, 2
GENOME, READ_LEN, COV = 5_000_000, 100, 30


@pytest.fixture(scope="module")
def c2():
    c = Cdbg(k=K, abundance_min=AMIN)
    n_reads = GENOME * COV // READ_LEN
    ptr, n = c.synth_reads(GENOME, READ_LEN, n_reads, 10000, 42)
    reads = c.copy_to_host(ptr, n)
    c.reset()
    c.push_device(ptr, n)
    c.count(); c.compact(); c.link()
    g = c.unitigs()
    yield c, reads, g
    c.close()


def test_c2_properties(c2):
    c, reads, g = c2
    st = g.stats
    k = K
    off = g.seq_off.astype(np.int64)
    ln = np.diff(off)
    assert st["n_unitigs"] == len(ln) and ln.min() >= k
    # every solid k-mer is spelled exactly once: sum over unitigs of (LN - k + 1) == n_solid, and KC sums to the total abundance
    assert int((ln - (k - 1)).sum()) == st["n_solid"]
    # the k-mers spelled by the unitigs are exactly the solid k-mer set (checksum of canonical k-mers, computed on the host)
    seq = g.seq
    codes = ((seq >> 1) & 3).astype(np.uint64)
    S = st["n_solid"]
    # rolling k-mers per unitig via cumulative trick would be heavy in numpy; sample 2000 unitigs instead and check them in the solid set
    km_gpu, ct_gpu = c.solid_kmers()
    order = np.argsort(km_gpu)
    km_sorted, ct_sorted = km_gpu[order], ct_gpu[order]
    assert len(np.unique(km_sorted)) == S
    assert int(ct_sorted.astype(np.uint64).sum()) == int(g.kc.sum())
    rng = np.random.default_rng(1)
    mask = (1 << (2 * k)) - 1
    for u in rng.choice(len(ln), size=2000, replace=False):
        s = codes[off[u]:off[u + 1]]
        fwd = 0; rc = 0; kc = 0
        for i, cd in enumerate(s.tolist()):
            fwd = ((fwd << 2) | cd) & mask
            rc = (rc >> 2) | ((cd ^ 2) << (2 * (k - 1)))
            if i >= k - 1:
                cm = min(fwd, rc)
                j = np.searchsorted(km_sorted, np.uint64(cm))
                assert j < S and int(km_sorted[j]) == cm, "unitig spells a non-solid k-mer"
                kc += int(ct_sorted[j])
        assert kc == int(g.kc[u])
    # links come in mirrored pairs: u -(s1,s2)-> v  <=>  v -(~s2,~s1)-> u
    L = g.links
    flip = {ord('+'): ord('-'), ord('-'): ord('+')}
    fwd_set = set(zip(L["from"].tolist(), L["from_sign"].tolist(), L["to"].tolist(), L["to_sign"].tolist()))
    for (a, s1, b, s2) in list(fwd_set)[:20000]:
        assert (b, flip[s2], a, flip[s1]) in fwd_set
    assert np.array_equal(np.diff(g.link_off.astype(np.int64)), np.bincount(L["from"], minlength=len(ln)))


def test_c2_full_graph_vs_oracle(c2):
    from oracle import oracle as orc
    c, reads, g = c2
    o = orc.run(reads, K, AMIN)
    assert o.stats["n_instances"] == g.stats["n_instances"]
    assert o.stats["n_distinct"] == g.stats["n_distinct"]
    assert o.stats["n_solid"] == g.stats["n_solid"]
    assert o.n_unitigs == g.n_unitigs and len(o.links) == len(g.links)
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), K)
    assert d == "", d


def test_c2_idempotent_and_deterministic_multiset(c2):
    """Running the whole path twice on the same context gives the same canonical graph (ids / strands may differ)."""
    c, reads, g = c2
    with Cdbg(k=K, abundance_min=AMIN) as c2_:
        g2 = c2_.build(reads)
    assert g2.n_unitigs == g.n_unitigs
    a = sorted(zip(np.diff(g.seq_off.astype(np.int64)).tolist(), g.kc.tolist()))
    b = sorted(zip(np.diff(g2.seq_off.astype(np.int64)).tolist(), g2.kc.tolist()))
    assert a == b


def test_c2_checksum_matches_the_committed_oracle_checksum(c2):
    """bench.py's N = 1 parity key compares with tests/golden/c2_checksum.json (the oracle's graph, made on CPU by
    tests/golden/make_c2_checksum.py with the numpy read generator): the device generator gives the same reads and the GPU graph
    the same checksum."""
    import json
    import os
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench
    from bcalm_b200.checksum import graph_checksum
    c, reads, g = c2
    n = 50_000 * (READ_LEN + 1)
    assert (reads[:n] == bench.synth_reads_numpy(GENOME, READ_LEN, 0, 50_000, 10000, 42)).all()
    with open(os.path.join(ROOT, "tests", "golden", "c2_checksum.json")) as f:
        gold = json.load(f)["scale_1"]
    cs = graph_checksum(g, device="cuda:0")
    assert {k_: cs[k_] for k_ in bench.CHECK_KEYS} == {k_: gold[k_] for k_ in bench.CHECK_KEYS}
    assert g.stats["n_solid"] == gold["n_solid"]


def test_c3_properties_on_one_gpu():
    """BASELINE.json configs[2] (250 Mbp genome, 50 M x 150 bp reads = 7.5 Gbases) at full size through size-independent properties:
    every solid k-mer is spelled once, KC sums to the solid abundances, every link overlaps by k-1, and a second build on a fresh
    context reproduces the checksum."""
    from bcalm_b200.checksum import graph_checksum
    G, L = 250_000_000, 150
    n_reads = G * 30 // L
    sums = []
    for _ in range(2):
        with Cdbg(k=K, abundance_min=AMIN) as c:
            ptr, n = c.synth_reads(G, L, n_reads, 10000, 43)
            c.push_device(ptr, n)
            st = c.count()
            km, ct = c.solid_kmers()
            assert len(km) == st["n_solid"]
            total_ab = int(ct.astype(np.uint64).sum())
            del km, ct
            c.compact(); c.link()
            g = c.unitigs(copy=False)
            cs = graph_checksum(g, device="cuda:0")
            assert cs["bad_overlaps"] == 0
            assert cs["total_nt"] - cs["n_unitigs"] * (K - 1) == st["n_solid"]
            assert cs["sum_kc"] == total_ab
            sums.append(cs)
    assert sums[0] == sums[1]
