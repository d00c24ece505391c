"""The sharded (multi-GPU) path run WITHOUT GPUs: N ranks = N host threads of this process, each driving its own context of
the CPU emulation build (tests/emu), with the NCCL calls of shard_host.inl served by an in-process stand-in
(tests/emu/nccl.h).  The orchestration, the exchange layouts (X1..X4, the all-gathers), the piece-level glue and the link
routing are the library's real code; only the transport and the execution of CUDA threads are emulated.  Complements
tests/test_shard_cpu.py (planning logic over a 2-process gloo group) and tests/test_gpu_sharded.py (the real thing, -m gpu).
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import threading

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from bcalm_b200 import api, parity  # noqa: E402
from oracle import oracle as orc  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the CPU emulation of the library")

with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
    GOLDEN = {c["name"]: c for c in json.load(f)}


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield path
    api.LIB_PATH, api._lib = saved


def sharded_build(reads, k, amin, world, **kw):
    """Round-robin the reads over `world` ranks (threads), build collectively, merge the shards."""
    uid = api.Cdbg.nccl_unique_id()
    out, err = [None] * world, [None] * world

    def work(rank):
        try:
            mine = reads[rank::world]
            buf = np.frombuffer(("\n".join(mine) + "\n").encode(), dtype=np.uint8) if mine else np.zeros(0, np.uint8)
            with api.Cdbg(k=k, abundance_min=amin, device=0, rank=rank, world_size=world, **kw) as cd:
                cd.shard_init(uid)
                out[rank] = cd.build(buf)
        except Exception as e:  # noqa: BLE001
            err[rank] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective"
    assert not any(err), err
    return api.merge_shards(out), out


CASES = sorted(n for n, c in GOLDEN.items() if c["unitigs_fa"] is not None)   # k <= 31 (minimizer-sharded) and k = 32..127 (hash-sharded, shard_wide_host.inl)


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", CASES)
def test_sharded_build_matches_reference_output(emu_lib, name, world):
    c = GOLDEN[name]
    m, shards = sharded_build(c["reads"], c["k"], c["abundance_min"], world)
    d = parity.diff(parity.graph_to_records(m), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d
    # global unitig ids are consecutive over the ranks, every rank owns a contiguous range
    first = 0
    for g in shards:
        assert g.first_unitig == first and g.n_unitigs_total == m.n_unitigs
        first += g.n_unitigs


def test_sharded_solid_kmer_count_equals_oracle(emu_lib):
    c = GOLDEN["reads_5_k25_a1"]
    m, shards = sharded_build(c["reads"], c["k"], c["abundance_min"], 4)
    o = orc.run(c["reads"], c["k"], c["abundance_min"])
    assert sum(g.stats["n_solid"] for g in shards) == o.stats["n_solid"]
    d = parity.diff(parity.graph_to_records(m), parity.graph_to_records(o), c["k"])
    assert d == "", d


# ---------------------------------------------------------------------------------------------------------------- ab:Z: on shards
AB_CASES = ["reads_0_k15_a2", "reads_3_k31_a2", "reads_5_k25_a1", "circular_test1", "lowcx_8_ACGT_k8_a1", "minitip_k21_a1", "even_k_reads_k20_a2", "circular_test2_polyA", "circular_genome_2_k21", "pow2_cycle_L256_k21", "lowcx_4_AT_k7_a2",
            "wide_1_k33_a2", "wide_4_k64_a2", "wide_8_k127_a2", "wide_lowcx_0_AT_k33"]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", [n for n in AB_CASES if n in GOLDEN])
def test_sharded_all_abundance_counts_match_the_single_gpu_lists(emu_lib, name, world):
    """-all-abundance-counts on shards: the per-k-mer lists travel with the pieces and are re-assembled (reversed pieces reversed)."""
    c = GOLDEN[name]
    with api.Cdbg(k=c["k"], abundance_min=c["abundance_min"], device=0, all_abundance_counts=True) as cd:
        one = cd.build(c["reads"])
    m, shards = sharded_build(c["reads"], c["k"], c["abundance_min"], world, all_abundance_counts=True)
    assert one.abund is not None and m.abund is not None and len(m.abund) == len(one.abund)
    r_one, r_m = parity.graph_to_records(one, with_ab=True), parity.graph_to_records(m, with_ab=True)
    assert all(sum(r["ab"]) == r["kc"] and len(r["ab"]) == r["ln"] - c["k"] + 1 for r in r_m)
    assert parity.diff_abundance_lists(r_m, r_one) == ""
    common = set(parity.abundance_lists(r_m)[0]) & set(parity.abundance_lists(r_one)[0])
    assert len(common) >= len(r_one) - 2 or name.startswith("circular")         # only circular unitigs may be spelled from another cut point
    d = parity.diff(parity.graph_to_records(m), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d


@pytest.mark.skipif(not orc.have_ref(), reason="needs the reference binary (oracle/_ref/bcalm)")
@pytest.mark.parametrize("k,world", [(63, 3), (127, 4), (33, 2)])
def test_sharded_multi_word_kmers_on_sequencing_like_reads_match_the_reference_binary(emu_lib, tmp_path, k, world):
    """60 kbp genome, 150 bp reads, 20x, 1 % errors, both strands: the hash-sharded k > 31 path against the reference itself (the C
    oracle stops at k = 31), per-k-mer abundance lists against the single-GPU path, and the shards must be balanced."""
    import random
    rng = random.Random(1000 + k)
    comp = str.maketrans("ACGT", "TGCA")
    genome = "".join(rng.choice("ACGT") for _ in range(60000))
    reads = []
    for _ in range(len(genome) * 20 // 150):
        s = rng.randrange(0, len(genome) - 150)
        r = genome[s:s + 150]
        if rng.random() < 0.5:
            r = r[::-1].translate(comp)
        reads.append("".join(c if rng.random() > 0.01 else rng.choice("ACGT") for c in r))
    fa = tmp_path / "x.fa"
    fa.write_text("".join(f">r\n{r}\n" for r in reads))
    ref = orc.run_ref(str(fa), k, 2, nb_cores=2, extra=["-verbose", "0"])[0]
    m, shards = sharded_build(reads, k, 2, world, all_abundance_counts=True)
    d = parity.diff(parity.graph_to_records(m), parity.parse_unitigs_fa(ref), k)
    assert d == "", d[:500]
    with api.Cdbg(k=k, abundance_min=2, device=0, all_abundance_counts=True) as c:
        one = c.build(reads)
    assert parity.diff_abundance_lists(parity.graph_to_records(m, with_ab=True), parity.graph_to_records(one, with_ab=True)) == ""
    solid = [s.stats["n_solid"] for s in shards]
    assert sum(solid) == one.stats["n_solid"] and max(solid) < 1.2 * min(solid) + 100


@pytest.mark.parametrize("k", [33, 64, 127])
def test_sharded_multi_word_kmers_degenerate_inputs(emu_lib, k):
    """Ranks without reads, reads without a single k-mer, no reads at all: the sharded k > 31 path must agree with one GPU (and, under the
    strict emulator, must not launch an empty grid on the way)."""
    import random
    rng = random.Random(5)
    r = "".join(rng.choice("ACGT") for _ in range(300))
    for reads in ([r], [r, r[10:200]], ["ACGT" * 5], []):
        with api.Cdbg(k=k, abundance_min=1, device=0) as c:
            one = c.build(reads if reads else "\n")
        for world in (2, 4):
            m, _ = sharded_build(reads, k, 1, world)
            d = parity.diff(parity.graph_to_records(m), parity.graph_to_records(one), k)
            assert d == "" and m.n_unitigs == one.n_unitigs, (len(reads), world, d[:300])


# ---------------------------------------------------------------------------------------------------------------- join variants on shards
@pytest.mark.parametrize("env", [{"CDBG_SHARD_K5": "l2"}, {"CDBG_K5S_FILL": "8"}, {"CDBG_K5S_FILL": "0.05"}, {"CDBG_SK_WAVES": "3"}],
                         ids=["l2_tables", "smem_forced_splits", "smem_tiny_fine_bins", "stage1_in_three_waves"])
@pytest.mark.parametrize("name", ["reads_3_k31_a2", "reads_0_k15_a2", "even_k_dense_k6", "lowcx_8_ACGT_k8_a1", "circular_test1", "pow2_cycle_L256_k21"])
def test_sharded_join_variants(emu_lib, monkeypatch, name, env):
    """The sharded join runs in shared-memory tables by default (k5cs_join); the L2-table kernels stay selectable, and table overflow
    (in-kernel splits) and many tiny fine bins give the same graph."""
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    c = GOLDEN[name]
    m, shards = sharded_build(c["reads"], c["k"], c["abundance_min"], 3)
    d = parity.diff(parity.graph_to_records(m), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    assert d == "", d
    if env.get("CDBG_K5S_FILL") == "8" and name == "reads_3_k31_a2":
        assert sum(g.stats["n_k5_splits"] for g in shards) > 0


def test_sharded_many_small_cycles_grow_the_edge_buffers(emu_lib, monkeypatch):
    """Cycles cut on a shard and cycles of pieces cut in the glue append their edges after the join has sized the buffers."""
    monkeypatch.setenv("CDBG_EDGE_MIN", "16")
    rng = np.random.default_rng(5)
    k = 13
    reads = []
    for _ in range(300):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, 19))
        reads.append(s + s[: k - 1])
    m, shards = sharded_build(reads, k, 1, 3)
    o = orc.run(reads, k, 1)
    d = parity.diff(parity.graph_to_records(m), parity.graph_to_records(o), k)
    assert d == "", d
