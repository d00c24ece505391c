"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from bcalm_b200 import parity
from bcalm_b200.api import Cdbg
from conftest import load_golden

pytestmark = pytest.mark.gpu

# every golden case gates (the even_k_* and ut_* cases passed on hardware: GPUTEST_r01, profiles/r02_gpu_tests_1gpu_start.txt)
CASES = load_golden()


def _name(c):
    return (c if isinstance(c, dict) else c.values[0])["name"]


def _records(g):
    return parity.graph_to_records(g)


@pytest.mark.parametrize("case", CASES, ids=[_name(c) for c in CASES])
def test_golden_vs_reference_output(case):
    """CUDA path == the real reference binary's unitigs.fa (committed golden), under the parity relation."""
    if case["unitigs_fa"] is None:
        pytest.skip("reference refused this input")
    k = case["k"]
    with Cdbg(k=k, abundance_min=case["abundance_min"], minimizer_size=case["minimizer_size"]) as c:
        g = c.build(case["reads"])
    d = parity.diff(_records(g), parity.parse_unitigs_fa(case["unitigs_fa"]), k)
    assert d == "", d


@pytest.mark.parametrize("case", CASES, ids=[_name(c) for c in CASES])
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


def _rand_reads(seed, glen, rl, n, alphabet="ACGT"):
    rng = np.random.default_rng(seed)
    g = "".join(rng.choice(list(alphabet), size=glen))
    return g, [g[s:s + rl] for s in rng.integers(0, glen - rl + 1, size=n)]


def test_push_reads_in_chunks_equals_single_push():
    """cdbg_push_reads appends (SortingCountAlgorithm consumes a bank of many files the same way): same graph as one push."""
    from oracle import oracle as orc
    _, reads = _rand_reads(11, 6000, 90, 1500)
    o = orc.run(reads, 25, 2)
    with Cdbg(k=25, abundance_min=2) as c:
        for i in range(0, len(reads), 400):
            c.push_reads(reads[i:i + 400])
        c.count(); c.compact(); c.link()
        g = c.unitigs()
    d = parity.diff(_records(g), parity.graph_to_records(o), 25)
    assert d == "", d


def test_empty_and_invalid_only_inputs():
    for reads in ([], ["NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"], ["ACGT"], ["\n\n\n"]):
        with Cdbg(k=21, abundance_min=1) as c:
            g = c.build(reads if reads else b"")
        assert g.n_unitigs == 0 and len(g.links) == 0 and g.stats["n_solid"] == 0


def test_abundance_max():
    """solid <=> abundance_min <= count <= abundance_max (CountProcessorSolidity.hpp:186-189)."""
    from oracle import oracle as orc
    _, reads = _rand_reads(12, 3000, 80, 1200)
    for amin, amax in [(2, 20), (1, 3), (5, 40)]:
        o = orc.run(reads, 21, amin, abundance_max=amax)
        with Cdbg(k=21, abundance_min=amin, abundance_max=amax) as c:
            g = c.build(reads)
        assert g.stats["n_solid"] == o.stats["n_solid"]
        d = parity.diff(_records(g), parity.graph_to_records(o), 21)
        assert d == "", d


@pytest.mark.parametrize("k", [15, 31, 47])
def test_long_sequences_across_many_tiles(k):
    """Genome-like input: a few very long lines (tile = 8192 positions, so every tile boundary and halo is crossed inside a sequence),
    with N runs and lower case sprinkled in."""
    from oracle import oracle as orc
    rng = np.random.default_rng(13 + k)
    g, _ = _rand_reads(14 + k, 120_000, 100, 1)
    g = list(g)
    for p in rng.integers(0, len(g), size=40):
        g[p] = "N"
    for p in rng.integers(0, len(g), size=200):
        g[p] = g[p].lower()
    g = "".join(g)
    reads = [g[:70_001], g[60_000:], parity.revcomp_str(g[20_000:45_017].upper())]
    o = orc.run_any(reads, k, 1)
    with Cdbg(k=k, abundance_min=1) as c:
        gg = c.build(reads)
    assert gg.stats["n_instances"] == o.stats["n_instances"] and gg.stats["n_solid"] == o.stats["n_solid"]
    d = parity.diff(_records(gg), parity.graph_to_records(o), k)
    assert d == "", d


def test_device_resident_input_unaligned_falls_back_to_copy():
    """CDBG_MEM_DEVICE input: used in place when 16-byte aligned, copied otherwise; both give the same graph."""
    from oracle import oracle as orc
    _, reads = _rand_reads(15, 5000, 100, 900)
    stream = np.frombuffer(("\n".join(reads) + "\n").encode(), dtype=np.uint8)
    o = orc.run(stream, 31, 2)
    import torch
    for shift in (0, 3):
        dev = torch.empty(stream.size + 64, dtype=torch.uint8, device="cuda")
        dev[shift:shift + stream.size] = torch.from_numpy(stream.copy()).cuda()
        with Cdbg(k=31, abundance_min=2) as c:
            c.push_device(dev.data_ptr() + shift, stream.size)
            torch.cuda.synchronize()
            c.count(); c.compact(); c.link()
            g = c.unitigs()
        d = parity.diff(_records(g), parity.graph_to_records(o), 31)
        assert d == "", d


def test_pinned_host_input_is_read_in_place_and_gives_the_same_graph():
    """cdbg_build with page-locked host reads: K1 reads the buffer over PCIe itself (stats.zero_copy), no separate H2D copy.
    (This path was timed and compared on the full C2 input in profiles/r01_k2_ab_2.txt / _3.txt.)"""
    from bcalm_b200.api import PinnedBuffer
    genome, read_len = 250_000, 100
    n_reads = genome * 30 // read_len
    with Cdbg(k=31, abundance_min=2) as c:
        ptr, n = c.synth_reads(genome, read_len, n_reads, 10000, 5)
        reads = c.copy_to_host(ptr, n)
        a = c.build(reads)                                   # pageable numpy memory: copied
        assert a.stats["zero_copy"] == 0
        with PinnedBuffer(n) as pb:
            pb.array[:] = reads
            b = c.build(pb.array)                            # pinned: read in place
            assert b.stats["zero_copy"] == 1
            assert b.stats["n_solid"] == a.stats["n_solid"] and b.stats["n_instances"] == a.stats["n_instances"]
            assert parity.diff(_records(a), _records(b), 31) == ""
