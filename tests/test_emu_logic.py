"""Kernel and host LOGIC of the CUDA library checked without a GPU.

tests/emu builds the unchanged library sources against a CPU stand-in for the CUDA runtime (every CUDA thread is an OS
thread, see tests/emu/cuda_runtime.h) and these tests push golden cases through that build via the normal ctypes mirror, for
the default configuration and for the variants / rare paths that the GPU suite reaches only through environment knobs:
the two K1 kernels behind the single-pass one (its overflow fallback), the split paths of the shared-memory K2 and K5, the
L2-table K2 and K5, match-based ranks, list-based pointer jumping, zero-copy input.  This is test infrastructure: it proves indexing, reservations and stage hand-offs, not performance and
not the GPU memory model -- the `-m gpu` parity tests remain the gate for the product.
"""
from __future__ import annotations

import json
import os
import shutil
import sys

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

KNOBS = ("CDBG_K1", "CDBG_K2", "CDBG_K2S_SLOTS", "CDBG_K2S_FILL", "CDBG_K2S_P1", "CDBG_K1_SLACK", "CDBG_ZEROCOPY", "CDBG_ZEROCOPY_MIN",
         "CDBG_EMU_PINNED", "CDBG_BIN_TARGET", "CDBG_JOIN_TARGET", "CDBG_PDL", "CDBG_K5", "CDBG_K5S_SLOTS", "CDBG_K5S_FILL", "CDBG_K5S_P1",
         "CDBG_RANK", "CDBG_K6", "CDBG_K2C", "CDBG_K1_THREADS", "CDBG_K2P", "CDBG_K2C_HASH",
         "CDBG_S1", "CDBG_SK_M", "CDBG_SK_SLOTS", "CDBG_SK_FILL", "CDBG_SK_P1", "CDBG_SK_DENSITY", "CDBG_SK_SLACK", "CDBG_SK2C", "CDBG_SK_RS", "CDBG_SK_KC", "CDBG_SK_WAVES", "CDBG_SK_INSERT", "CDBG_SK_THREADS", "CDBG_SK_WCAP")


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield path
    api.LIB_PATH, api._lib = saved


@pytest.fixture()
def knobs(monkeypatch):
    for k in KNOBS:
        monkeypatch.delenv(k, raising=False)

    def set_(**kw):
        for k, v in kw.items():
            monkeypatch.setenv(k, str(v))
    return set_


def run_case(case, **kw):
    with api.Cdbg(k=case["k"], abundance_min=case["abundance_min"], device=0, **kw) as c:
        return c.build(case["reads"])


def check_against_golden(g, case):
    d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(case["unitigs_fa"]), case["k"])
    assert d == "", d


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_default_configuration_matches_reference_output(emu_lib, knobs, name):
    g = run_case(GOLDEN[name])
    check_against_golden(g, GOLDEN[name])


KMER = dict(CDBG_S1="kmer")    # stage 1 on k-mer instances (k1_tma.cuh / k2_smem.cuh): the round-1 default, now behind CDBG_S1=kmer
VARIANTS = {
    # stage 1 on super-k-mers (sk.cuh, the default)
    "sk_m_4": dict(CDBG_SK_M=4),
    "sk_m_16": dict(CDBG_SK_M=16),
    "sk_many_coarse_bins": dict(CDBG_SK_P1=37),
    "sk_tiny_fine_bins": dict(CDBG_SK_FILL=0.05),
    "sk_forced_splits": dict(CDBG_SK_FILL=8),
    "sk_slots_8192": dict(CDBG_SK_SLOTS=8192, CDBG_SK_FILL=0.2),
    "sk_mul32_slot_hash": dict(CDBG_K2C_HASH="mul32"),
    "sk_three_waves": dict(CDBG_SK_WAVES=3),
    "sk_waves_and_exact_regions": dict(CDBG_SK_WAVES=2, CDBG_SK_DENSITY=0.001, CDBG_SK_SLACK=0, CDBG_SK_P1=5),
    "sk_waves_zero_copy": dict(CDBG_SK_WAVES=4, CDBG_EMU_PINNED=1, CDBG_ZEROCOPY_MIN=0),
    "sk_two_phase_insert": dict(CDBG_SK_INSERT="two"),
    "sk_two_phase_forced_splits": dict(CDBG_SK_INSERT="two", CDBG_SK_FILL=8),
    "sk_two_phase_8192_runtime_k": dict(CDBG_SK_INSERT="two", CDBG_SK_SLOTS=8192, CDBG_SK_KC=0, CDBG_K2C_HASH="mix64"),
    "sk_small_staging_area": dict(CDBG_SK_WCAP=128),
    "sk_512_threads_per_table": dict(CDBG_SK_THREADS=512),
    "sk_512_threads_forced_splits_loop": dict(CDBG_SK_THREADS=512, CDBG_SK_FILL=8, CDBG_SK_INSERT="loop"),
    "sk_direct_count": dict(CDBG_SK2C="direct"),
    "sk_dedup_count": dict(CDBG_SK2C="dedup"),
    "sk_dedup_small_record_table": dict(CDBG_SK2C="dedup", CDBG_SK_RS=512, CDBG_SK_FILL=3),
    "sk_dedup_forced_splits": dict(CDBG_SK2C="dedup", CDBG_SK_FILL=8),
    "sk_dedup_runtime_k_8192": dict(CDBG_SK2C="dedup", CDBG_SK_SLOTS=8192, CDBG_SK_KC=0, CDBG_SK_RS=2048),
    "sk_exact_regions_after_overflow": dict(CDBG_SK_DENSITY=0.001, CDBG_SK_SLACK=0, CDBG_SK_P1=5),
    "sk_zero_copy_input": dict(CDBG_EMU_PINNED=1, CDBG_ZEROCOPY_MIN=0),
    "sk_zero_copy_overflow": dict(CDBG_EMU_PINNED=1, CDBG_ZEROCOPY_MIN=0, CDBG_SK_DENSITY=0.001, CDBG_SK_SLACK=0),
    # stage 1 on k-mer instances
    "k1_single_pass": dict(**KMER),
    "k1_twopass_tma": dict(CDBG_K1="twopass", **KMER),
    "k1_classic_ldg": dict(CDBG_K1="classic", **KMER),
    "k2_l2_tables": dict(CDBG_K2="l2", **KMER),
    "k2_l2_many_bins": dict(CDBG_K2="l2", CDBG_BIN_TARGET=3000, **KMER),
    "k2_smem_many_fine_bins": dict(CDBG_K2S_SLOTS=4096, CDBG_K2S_FILL=0.05, **KMER),
    "k2_smem_16384": dict(CDBG_K2S_SLOTS=16384, CDBG_K2S_FILL=0.2, **KMER),
    "k2_smem_forced_splits": dict(CDBG_K2S_SLOTS=4096, CDBG_K2S_FILL=8, **KMER),
    "k1_many_coarse_bins": dict(CDBG_K2S_P1=37, **KMER),
    "zero_copy_input": dict(CDBG_EMU_PINNED=1, CDBG_ZEROCOPY_MIN=0, **KMER),
    "ranks_by_match": dict(CDBG_RANK="match", CDBG_K2S_P1=9, **KMER),
    "k2c_flat_insert_loop": dict(CDBG_K2C="flat", **KMER),
    "k2c_mul32_slot_hash": dict(CDBG_K2C_HASH="mul32", **KMER),
    "k2c_mul32_flat_forced_splits": dict(CDBG_K2C_HASH="mul32", CDBG_K2C="flat", CDBG_K2S_SLOTS=4096, CDBG_K2S_FILL=8, **KMER),
    "k2c_flat_forced_splits": dict(CDBG_K2C="flat", CDBG_K2S_SLOTS=4096, CDBG_K2S_FILL=8, **KMER),
    "k1_512_threads": dict(CDBG_K1_THREADS=512, CDBG_K2S_P1=5, **KMER),
    "k2p_regions": dict(CDBG_K2P="regions", CDBG_K2S_P1=3, **KMER),
    "k2p_regions_tiny_bins": dict(CDBG_K2P="regions", CDBG_K2S_FILL=0.05, CDBG_K2S_SLOTS=4096, **KMER),
    # stages 2-4
    "no_pdl": dict(CDBG_PDL=0),
    "k5_l2_tables": dict(CDBG_K5="l2"),
    "k5_l2_many_bins": dict(CDBG_K5="l2", CDBG_JOIN_TARGET=500),
    "k5_smem_many_fine_bins": dict(CDBG_K5S_FILL=0.05, CDBG_K5S_P1=7),
    "k5_smem_forced_splits": dict(CDBG_K5S_FILL=8, CDBG_K5S_SLOTS=4096),
    "k6_over_lists": dict(CDBG_K6="list"),
    "k6_two_level": dict(CDBG_K6="local"),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("name", ["circular_test1", "circular_test2_polyA", "lowcx_3_AT_k9_a1", "lowcx_8_ACGT_k8_a1", "palindrome_k_even",
                                  "pow2_cycle_L256_k21", "reads_0_k15_a2", "reads_3_k31_a2", "reads_5_k25_a1"])
def test_variants_match_reference_output(emu_lib, knobs, variant, name):
    knobs(**VARIANTS[variant])
    g = run_case(GOLDEN[name])
    check_against_golden(g, GOLDEN[name])
    if variant in ("zero_copy_input", "sk_zero_copy_input", "sk_zero_copy_overflow"):
        assert g.stats["zero_copy"] == 1
    if variant == "sk_three_waves":
        assert g.stats["sk_waves"] == 3
    if variant.startswith("sk_"):
        assert g.stats["sk_m"] > 0 and g.stats["n_superkmers"] > 0
    elif variant.startswith(("k1_", "k2")):
        assert g.stats["sk_m"] == 0
    if variant in ("sk_exact_regions_after_overflow", "sk_zero_copy_overflow") and name.startswith("reads_"):
        assert g.stats["k1_fallback"] == 1
    if variant in ("k2_smem_forced_splits", "sk_forced_splits", "sk_dedup_forced_splits") and name.startswith("reads_"):
        assert g.stats["n_k2_splits"] > 0
    if variant == "k5_smem_forced_splits" and name.startswith("reads_"):
        assert g.stats["n_k5_splits"] > 0


def _synthetic_reads(genome_len, read_len, coverage, err, seed):
    """Sequencing-like reads (both strands, substitution errors) over a random genome, as in bench.py's workload."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len)
    n = genome_len * coverage // read_len
    starts = rng.integers(0, genome_len - read_len + 1, n)
    comp = np.array([3, 2, 1, 0])
    reads = []
    for st_ in starts:
        r = genome[st_:st_ + read_len].copy()
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        e = rng.random(read_len) < err
        r[e] = (r[e] + rng.integers(1, 4, int(e.sum()))) % 4
        reads.append("".join("ACGT"[i] for i in r))
    return reads


@pytest.mark.parametrize("variant", ["default", "sk_forced_splits", "sk_exact_regions_after_overflow", "sk_m_16", "sk_direct_count", "sk_dedup_count",
                                     "sk_dedup_small_record_table", "sk_dedup_forced_splits", "k6_two_level", "sk_512_threads_per_table", "sk_two_phase_insert", "sk_two_phase_forced_splits", "sk_three_waves", "sk_waves_and_exact_regions", "k1_single_pass", "k1_512_threads", "k2_l2_tables",
                                     "k5_l2_tables", "k1_twopass_tma", "k2_smem_forced_splits", "k5_smem_forced_splits"])
def test_larger_synthetic_read_set_matches_oracle(emu_lib, knobs, variant):
    """40 kbp genome, 100 bp reads at 30x with 1 % errors (1.2 Mbases): several tiles, coarse and fine bins, real branching."""
    if variant != "default":
        knobs(**VARIANTS[variant])
    reads = _synthetic_reads(40_000, 100, 30, 0.01, 11)
    with api.Cdbg(k=31, abundance_min=2, device=0) as c:
        g = c.build(reads)
    o = orc.run(reads, 31, 2)
    assert g.stats["n_solid"] == o.stats["n_solid"]
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), 31)
    assert d == "", d


def test_truly_concurrent_engine_smoke():
    """The default emulator engine runs the CUDA threads of a block as fibers of one OS thread.  Run a few cases once with one OS
    thread per CUDA thread as well (real races on the shared-memory and global atomics), in a subprocess: the engine is chosen
    at the first launch of a process."""
    import subprocess
    code = (
        "import json, os, sys\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, 'tests', 'emu'))\n"
        "import build_emu\n"
        "from bcalm_b200 import api, parity\n"
        "api.LIB_PATH = build_emu.build()\n"
        f"cases = {{c['name']: c for c in json.load(open(os.path.join({ROOT!r}, 'tests', 'golden', 'golden.json')))}}\n"
        "for n in ('minitip_k21_a1', 'circular_test1', 'reads_0_k15_a2'):\n"
        "    c = cases[n]\n"
        "    with api.Cdbg(k=c['k'], abundance_min=c['abundance_min'], device=0) as x:\n"
        "        g = x.build(c['reads'])\n"
        "    assert parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c['unitigs_fa']), c['k']) == '', n\n"
        "print('ok')\n")
    env = dict(os.environ, CDBG_EMU_ENGINE="threads")
    for k_ in KNOBS:
        env.pop(k_, None)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def _repeats_input():
    """One heavily repeated k-mer family (poly-A reads) next to ordinary reads: the repeated k-mer lands in one bin."""
    rng = np.random.default_rng(7)
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, 600))
    reads = ["A" * 100] * 300 + [genome[i:i + 80] for i in range(0, 520, 3)] * 2
    return reads


def test_k1_single_pass_overflow_falls_back_to_exact_two_pass(emu_lib, knobs):
    reads = _repeats_input()
    knobs(CDBG_K2S_P1=8, CDBG_K1_SLACK=0, CDBG_S1="kmer")
    with api.Cdbg(k=21, abundance_min=2, device=0) as c:
        g = c.build(reads)
    assert g.stats["k1_fallback"] == 1
    o = orc.run(reads, 21, 2)
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), 21)
    assert d == "", d


def test_k2_region_partition_overflow_falls_back_to_exact_offsets(emu_lib, knobs):
    reads = _repeats_input()
    knobs(CDBG_K2P="regions", CDBG_K2S_FILL=0.05, CDBG_K2S_SLOTS=4096, CDBG_K2S_P1=2, CDBG_S1="kmer")
    with api.Cdbg(k=21, abundance_min=2, device=0) as c:
        g = c.build(reads)
    assert g.stats["k2p_fallback"] == 1
    o = orc.run(reads, 21, 2)
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), 21)
    assert d == "", d


def test_zero_copy_overflow_falls_back_with_a_copy(emu_lib, knobs):
    reads = _repeats_input()
    knobs(CDBG_K2S_P1=8, CDBG_K1_SLACK=0, CDBG_EMU_PINNED=1, CDBG_ZEROCOPY_MIN=0, CDBG_S1="kmer")
    with api.Cdbg(k=21, abundance_min=2, device=0) as c:
        g = c.build(reads)
    assert g.stats["k1_fallback"] == 1 and g.stats["zero_copy"] == 1
    o = orc.run(reads, 21, 2)
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), 21)
    assert d == "", d


def _many_runs_input(k):
    """Long reads whose minimizer window is only two m-mers wide (k small, m = k-1 is forced by CDBG_SK_M): a run is ~1.5 k-mers long,
    so a tile of 8192 positions holds several thousand runs -- more than the tile's shared-memory run list (2048) -- and the poly-A
    stretches are single runs of thousands of k-mers, cut into many records -- more than the staging area holds."""
    rng = np.random.default_rng(5)
    reads = []
    for _ in range(6):
        body = "".join("ACGT"[i] for i in rng.integers(0, 4, 7000))
        reads.append(body + "A" * 3000 + body[:500])
    return reads


@pytest.mark.parametrize("k,m", [(5, 4), (7, 6), (9, 3), (31, 16)])
def test_superkmer_extraction_beyond_the_per_tile_capacities(emu_lib, knobs, k, m):
    reads = _many_runs_input(k)
    knobs(CDBG_SK_M=m)
    with api.Cdbg(k=k, abundance_min=1, device=0) as c:
        g = c.build(reads)
    o = orc.run(reads, k, 1)
    assert g.stats["n_instances"] == o.stats["n_instances"] and g.stats["n_solid"] == o.stats["n_solid"]
    if k - m == 1:
        assert g.stats["n_superkmers"] > 2048 * 6          # (the direct path was taken)
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
    assert d == "", d


def test_superkmer_record_layout_is_what_the_header_says(emu_lib, knobs):
    """Every k-mer instance is in exactly one record: the instance count equals the sum of the records' k-mer counts, and a read
    with no minimizer change (poly-A) becomes ceil(n / (53 - k)) records."""
    k = 21
    with api.Cdbg(k=k, abundance_min=1, device=0) as c:
        c.push_reads(["A" * 200])
        st = c.count()
    n = 200 - k + 1
    assert st["n_instances"] == n and st["n_superkmers"] == -(-n // (53 - k)) and st["n_solid"] == 1


def test_stage_one_in_isolation_matches_oracle_counts(emu_lib, knobs):
    case = GOLDEN["reads_0_k15_a2"]
    with api.Cdbg(k=case["k"], abundance_min=case["abundance_min"], device=0) as c:
        c.push_reads(case["reads"])
        st = c.count()
        km, ct = c.solid_kmers()
    o = orc.run(case["reads"], case["k"], case["abundance_min"])
    assert st["n_solid"] == len(km)
    order = np.argsort(km)
    assert int(ct.sum()) == int(np.asarray(o.kc).sum())
    assert len(np.unique(km)) == len(km) and np.all(km[order][1:] > km[order][:-1])


def test_counted_kmers_hand_off_through_the_abi(emu_lib, knobs):
    """cdbg_set_solid_kmers: stage 1 result of one context (half of the k-mers flipped to the other strand) -> stages 2-4 of another."""
    case = GOLDEN["reads_3_k31_a2"]
    k = case["k"]
    with api.Cdbg(k=k, abundance_min=case["abundance_min"], device=0) as a:
        a.push_reads(case["reads"])
        a.count()
        km, ct = a.solid_kmers()
    rc = np.zeros_like(km)
    for i in range(k):
        rc = (rc << np.uint64(2)) | (((km >> np.uint64(2 * i)) & np.uint64(3)) ^ np.uint64(2))
    mixed = np.where(np.arange(len(km)) % 2 == 0, rc, km)
    with api.Cdbg(k=k, abundance_min=1, device=0) as b:
        b.set_solid_kmers(mixed, ct)
        b.compact()
        b.link()
        g = b.unitigs()
        with pytest.raises(api.CdbgError):
            b.set_solid_kmers(mixed, ct)          # not on a context that already holds results
    check_against_golden(g, case)
    with api.Cdbg(k=63, abundance_min=1, device=0) as w, pytest.raises(ValueError):
        w.set_solid_kmers(km, ct)                 # k = 63 takes (n, 2) words


@pytest.mark.parametrize("name", ["wide_0_k32_a2", "wide_3_k63_a2", "wide_5_k65_a2", "wide_8_k127_a2"])
def test_counted_kmers_hand_off_with_multi_word_kmers(emu_lib, knobs, name):
    """The same hand-off for k = 32..127: (n, W) little-endian words, half of them handed over as the other strand."""
    case = GOLDEN[name]
    k = case["k"]
    with api.Cdbg(k=k, abundance_min=case["abundance_min"], device=0) as a:
        a.push_reads(case["reads"])
        a.count()
        km, ct = a.solid_kmers()
    W = km.shape[1]
    mixed = km.copy()
    for i in range(0, len(km), 2):
        x = sum(int(km[i, w]) << (64 * w) for w in range(W))
        r = 0
        for j in range(k):
            r = (r << 2) | (((x >> (2 * j)) & 3) ^ 2)
        mixed[i] = [(r >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(W)]
    assert not np.array_equal(mixed, km)
    with api.Cdbg(k=k, abundance_min=1, device=0) as b:
        b.set_solid_kmers(mixed, ct)
        b.compact()
        b.link()
        g = b.unitigs()
    check_against_golden(g, case)


def test_repeated_builds_on_one_context_are_identical(emu_lib, knobs):
    case = GOLDEN["reads_6_k31_a1"]
    with api.Cdbg(k=case["k"], abundance_min=case["abundance_min"], device=0) as c:
        a = c.build(case["reads"])
        b = c.build(case["reads"])
    assert parity.diff(parity.graph_to_records(a), parity.graph_to_records(b), case["k"]) == ""
    check_against_golden(b, case)


def test_reference_known_answers_through_the_cuda_path(emu_lib, knobs):
    """The known-answer vectors of the reference's own test tree that pin this path, through the emulated CUDA stage 1
    and the full path: TestDSK.cpp:147-175 (DSK_check1) and the example/ + test/ fixtures quoted in SURVEY.md."""
    s1 = "GATCCTCCCCAGGCCCCTACACCCAAT"

    def n_solid(reads, k, a):
        with api.Cdbg(k=k, abundance_min=a, device=0) as c:
            c.push_reads(reads)
            return c.count()["n_solid"]
    assert n_solid([s1], 27, 1) == 1
    assert n_solid([s1], 26, 1) == 2
    assert n_solid([s1, s1], 27, 2) == 1
    assert n_solid([s1, s1], 26, 2) == 2
    assert n_solid([s1, s1], 27, 3) == 0
    # example/tiny_read.fa: k=21 -> one unitig of 25 nt, KC 5; k=13 -> KC 13
    tiny = GOLDEN["tiny_read_k21"]["reads"]
    with api.Cdbg(k=21, abundance_min=1, device=0) as c:
        g = c.build(tiny)
    assert g.n_unitigs == 1 and len(g.seq) == 25 and int(g.kc[0]) == 5 and len(g.links) == 0
    with api.Cdbg(k=13, abundance_min=1, device=0) as c:
        g = c.build(tiny)
    assert g.n_unitigs == 1 and int(g.kc[0]) == 13


def test_second_push_after_a_buffer_used_in_place_is_separated_by_a_break(emu_lib, knobs):
    """a first push of an aligned device buffer is borrowed in place; a later push must not glue its first read to the
    borrowed buffer's last one (the borrowed bytes need not end in a break byte)."""
    k = 15
    with api.Cdbg(k=k, abundance_min=1, device=0) as c:
        ptr, n = c.synth_reads(3000, 60, 40, 10000, 5)
        a = c.copy_to_host(ptr, n)[: n - 1]                     # drop the final '\n': the buffer now ends in a nucleotide
        b = np.frombuffer(b"ACGTTGCATGCCGATAGCTAGGATCCGATTTACGCAGT\n", dtype=np.uint8)
        c.push_device(ptr, n - 1)
        c.push_reads(b)
        assert c.pushed_reads() == a.tobytes() + b"\n" + b.tobytes() + b"\n"
        c.count(); c.compact(); c.link()
        g = c.unitigs()
    o = orc.run(np.concatenate([a, np.frombuffer(b"\n", np.uint8), b]), k, 1)
    assert parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k) == ""


def _many_cycles(n_cycles, length, k, seed):
    """Reads that each spell one isolated cycle of `length` k-mers (the circular string plus its first k-1 characters)."""
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_cycles):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, length))
        reads.append(s + s[: k - 1])
    return reads


def test_many_small_cycles_grow_the_edge_buffers(emu_lib, knobs, monkeypatch):
    """the edge buffer kept room for ~1000 cut cycles and failed beyond.  Hundreds of isolated cycles with the buffers
    started at 16 entries (CDBG_EDGE_MIN): the join's retry and the cycle cutter's growth path must produce the oracle's graph."""
    monkeypatch.setenv("CDBG_EDGE_MIN", "16")
    k = 13
    reads = _many_cycles(400, 17, k, 3)
    with api.Cdbg(k=k, abundance_min=1, device=0) as c:
        g = c.build(reads)
    o = orc.run(reads, k, 1)
    assert g.stats["n_cycles"] >= 300
    d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
    assert d == "", d
