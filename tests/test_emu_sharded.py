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


def sharded_build(reads, k, amin, world):
    """Round-robin the reads over `world` ranks (threads), build collectively, merge the shards."""
    uid = api.Cdbg.nccl_unique_id()
    out, err = [None] * world, [None] * world

    def work(rank):
        try:
            mine = reads[rank::world]
            buf = np.frombuffer(("\n".join(mine) + "\n").encode(), dtype=np.uint8) if mine else np.zeros(0, np.uint8)
            with api.Cdbg(k=k, abundance_min=amin, device=0, rank=rank, world_size=world) as cd:
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


CASES = sorted(n for n, c in GOLDEN.items() if c["k"] <= 31)   # (k > 31 is single-GPU only)


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
