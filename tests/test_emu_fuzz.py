"""Randomised parity sweep of the CUDA path (under the CPU emulator) against the oracle: small k (even and odd), low-complexity
alphabets, reads of exactly k / k+1 nucleotides, N's and lower case, byte streams that end within a few bytes of a tile edge,
and a randomly chosen kernel configuration per case (default, the L2-table and two-pass variants, the opt-in variants, tiny
fine bins, forced splits, zero-copy input, and the sharded path over 2-3 emulated ranks).  Seeded, so a failure reproduces.
This sweep is what found the 5-ends-per-junction-side case of even k (tests/golden: even_k_dense_*)."""
from __future__ import annotations

import os
import random
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

from bcalm_b200 import api, parity  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_emu_sharded import sharded_build  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the CPU emulation of the library")

CONFIGS = [{}, {"CDBG_K1": "twopass"}, {"CDBG_K1": "classic"}, {"CDBG_K2": "l2"}, {"CDBG_K5": "l2"}, {"CDBG_K2C": "flat"}, {"CDBG_K2C_HASH": "mul32"}, {"CDBG_K1_THREADS": "512"},
           {"CDBG_K6": "list"}, {"CDBG_K2P": "regions"}, {"CDBG_K2P": "regions", "CDBG_K2S_FILL": "0.1", "CDBG_K2S_SLOTS": "4096"}, {"CDBG_RANK": "match", "CDBG_K2S_P1": "6"}, {"CDBG_EMU_PINNED": "1", "CDBG_ZEROCOPY_MIN": "0"},
           {"CDBG_K2S_P1": "7", "CDBG_K2S_FILL": "0.05", "CDBG_K2S_SLOTS": "4096"}, {"CDBG_K5S_P1": "5", "CDBG_K5S_FILL": "0.05"},
           {"CDBG_K2S_FILL": "8", "CDBG_K2S_SLOTS": "4096", "CDBG_K5S_FILL": "8"}, {"shards": 2}, {"shards": 3},
           {"shards": 3, "CDBG_SHARD_K2": "smem", "CDBG_K2S_FILL": "0.05", "CDBG_K2S_SLOTS": "4096"}, {"shards": 2, "CDBG_SHARD_K2": "smem"},
           {"shards": 2, "CDBG_SHARD_K1": "roll"}, {"shards": 4, "CDBG_SHARD_K1": "roll", "CDBG_SHARD_K2": "smem"}]
KNOBS = sorted({k for c in CONFIGS for k in c if k.startswith("CDBG_")})
COMP = str.maketrans("ACGTacgt", "TGCAtgca")


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield path
    api.LIB_PATH, api._lib = saved


def make_case(rng: random.Random, wide: bool = False):
    k = rng.choice([32, 33, 41, 48, 63, 64, 65, 96, 127]) if wide else rng.choice([4, 5, 6, 7, 8, 11, 15, 20, 21, 25, 31, 31, 31])
    alpha = rng.choice(["ACGT", "ACGT", "AT", "ACG", "ACGTN", "ACGTacgt"])
    glen = rng.choice([50, 300, 2000, 9000, 20000])
    letters = "ACGT" if alpha in ("ACGTN", "ACGTacgt") else alpha
    genome = "".join(rng.choice(letters) for _ in range(glen))
    L = rng.choice([k, k + 1, 40, 100, 150, 251])
    cov = rng.choice([1, 3, 8])
    reads = []
    for _ in range(max(1, glen * cov // L)):
        s = rng.randrange(0, max(1, glen - L + 1))
        r = genome[s:s + L]
        if rng.random() < 0.5:
            r = r[::-1].translate(COMP)
        if alpha == "ACGTN" and rng.random() < 0.3:
            p = rng.randrange(len(r)); r = r[:p] + "N" + r[p + 1:]
        if alpha == "ACGTacgt" and rng.random() < 0.5:
            r = r.lower()
        if rng.random() < 0.02 and len(r) > 2:
            p = rng.randrange(len(r)); r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        reads.append(r)
    if rng.random() < 0.4:   # end the byte stream within +-40 bytes of a tile edge (8192-byte tiles, 32-byte halo)
        tot = sum(len(r) + 1 for r in reads)
        target = (tot // 8192 + 1) * 8192 + rng.randint(-40, 40)
        while tot < target - 1:
            r = genome[:min(len(genome), target - 1 - tot)] or "A"
            reads.append(r); tot += len(r) + 1
    return k, rng.choice([1, 1, 2, 3]), reads, rng.choice(CONFIGS)


@pytest.mark.parametrize("seed", range(8))
def test_random_inputs_and_configurations_match_the_oracle(emu_lib, monkeypatch, seed):
    rng = random.Random(1000 + seed)
    for it in range(40):
        k, amin, reads, cfg = make_case(rng)
        for name in KNOBS:
            monkeypatch.delenv(name, raising=False)
        for name, v in cfg.items():
            if name.startswith("CDBG_"):
                monkeypatch.setenv(name, v)
        o = orc.run(reads, k, amin)
        if "shards" in cfg:
            g, parts = sharded_build(reads, k, amin, cfg["shards"])
            n_solid = sum(p.stats["n_solid"] for p in parts)
        else:
            with api.Cdbg(k=k, abundance_min=amin, device=0) as c:
                g = c.build(reads)
            n_solid = g.stats["n_solid"]
            assert g.stats["n_instances"] == o.stats["n_instances"] and g.stats["n_distinct"] == o.stats["n_distinct"], (seed, it, k, cfg)
        assert n_solid == o.stats["n_solid"], (seed, it, k, amin, cfg)
        d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
        assert d == "", (seed, it, k, amin, cfg, d[:400])


ALL_KNOBS = {"CDBG_K1": ["single", "twopass", "classic"], "CDBG_K1_THREADS": ["256", "512"], "CDBG_K2": ["smem", "l2"], "CDBG_K2P": ["exact", "regions"],
             "CDBG_K2C": ["nested", "flat"], "CDBG_K2C_HASH": ["mix64", "mul32"], "CDBG_K5": ["smem", "l2"], "CDBG_K6": ["full", "list"], "CDBG_RANK": ["atomic", "match"],
             "CDBG_K2S_SLOTS": ["4096", "8192", "16384"], "CDBG_K2S_FILL": ["0.05", "0.7", "1.5", "8"], "CDBG_K2S_P1": ["1", "3", "11"],
             "CDBG_K5S_FILL": ["0.05", "1.0", "8"], "CDBG_K5S_SLOTS": ["2048", "4096"], "CDBG_K5S_P1": ["1", "4"], "CDBG_K1_SLACK": ["0", "16384"],
             "CDBG_EMU_PINNED": ["0", "1"], "CDBG_ZEROCOPY_MIN": ["0"], "CDBG_PDL": ["0", "1"], "CDBG_SHARD_K1": ["classic", "roll"],
             "CDBG_SHARD_K2": ["l2", "smem"], "CDBG_BIN_TARGET": ["3000", "5000000"], "CDBG_JOIN_TARGET": ["700", "500000"]}


def _reference_output(reads, k, amin, tmp_path):
    fa = tmp_path / "fuzz.fa"
    fa.write_text("".join(f">r{i}\n{r}\n" for i, r in enumerate(reads)))
    try:
        return orc.run_ref(str(fa), k, amin, nb_cores=1, extra=["-verbose", "0"], timeout=120)[0]
    except Exception:  # noqa: BLE001 -- the reference itself gives up on some degenerate inputs (nothing solid, ...)
        return None


def _single_build(reads, k, amin):
    with api.Cdbg(k=k, abundance_min=amin, device=0) as c:
        return c.build(reads)


@pytest.mark.parametrize("seed", range(3))
def test_random_combinations_of_all_knobs(emu_lib, monkeypatch, tmp_path, seed):
    """Every knob is meant to be output-neutral on its own AND together with any other: random subsets, single GPU and sharded."""
    rng = random.Random(5000 + seed)
    for it in range(40):
        wide = rng.random() < 0.2                     # multi-word k-mers too (single GPU and hash-sharded); checked against the
        k, amin, reads, _ = make_case(rng, wide=wide)  # reference binary itself (the C oracle is k <= 31), skipped when it is absent
        env = {name: rng.choice(vals) for name, vals in ALL_KNOBS.items() if rng.random() < 0.5}
        for name in ALL_KNOBS:
            monkeypatch.delenv(name, raising=False)
        for name, v in env.items():
            monkeypatch.setenv(name, v)
        shards = rng.choice([0, 0, 2, 3])
        ab = rng.random() < 0.4 and not wide         # -all-abundance-counts: per-k-mer lists, also on shards
        if wide:
            if not orc.have_ref():
                continue
            ref_text = _reference_output(reads, k, amin, tmp_path)
            if ref_text is None:
                continue
            g = sharded_build(reads, k, amin, shards)[0] if shards else _single_build(reads, k, amin)
            d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(ref_text), k)
            assert d == "", (seed, it, k, amin, shards, d[:400])
            continue
        o = orc.run(reads, k, amin)
        if shards:
            g, parts = sharded_build(reads, k, amin, shards, all_abundance_counts=ab)
            n_solid = sum(p.stats["n_solid"] for p in parts)
        else:
            with api.Cdbg(k=k, abundance_min=amin, device=0, all_abundance_counts=ab) as c:
                g = c.build(reads)
            n_solid = g.stats["n_solid"]
        assert n_solid == o.stats["n_solid"], (seed, it, k, amin, shards, env)
        d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
        assert d == "", (seed, it, k, amin, shards, env, d[:400])
        if ab:
            d = parity.diff_abundance_lists(parity.graph_to_records(g, with_ab=True), parity.graph_to_records(o, with_ab=True))
            assert d == "", (seed, it, k, amin, shards, env, d)
