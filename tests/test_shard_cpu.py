"""CPU tests of the sharded (N > 1) host logic with a world_size-2 gloo process group: the ownership function and the
destination-major exchange planning are the library's own code (cdbg_owner_of / cdbg_plan_shard_exchange, no GPU needed); the
k-mers and minimizers come from the oracle, the transport is gloo instead of NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, M, P_LOCAL = 21, 8, 5


def _mix32(x):
    x = (x ^ 0x68E31DA4) & 0xFFFFFFFF
    x ^= x >> 15; x = (x * 0x2C1B3C6D) & 0xFFFFFFFF
    x ^= x >> 12; x = (x * 0x297A2D39) & 0xFFFFFFFF
    x ^= x >> 15
    return x


def test_owner_matches_oracle_minimizer():
    """owner = hash(mmer_rank(minimizer)) * world >> 32, with the oracle's independent minimizer (orc_minimizer/orc_mmer_rank)."""
    from bcalm_b200.api import owner_of
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    L = orc.lib()
    for n, m in [(31, 10), (30, 10), (21, 8), (20, 8), (9, 5), (8, 8), (5, 4)]:
        for _ in range(300):
            x = int(rng.integers(0, 1 << (2 * n), dtype=np.uint64))
            mini = orc.minimizer(x, n, m)
            rank = int(L.orc_mmer_rank(mini))
            for world in (1, 2, 3, 8):
                assert owner_of(x, n, m, world) == (_mix32(rank) * world) >> 32
            # strand symmetry: an n-mer and its reverse complement have the same owner
            assert owner_of(orc.revcomp(x, n), n, m, 8) == owner_of(x, n, m, 8)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bcalm_b200.api import owner_of, plan_shard_exchange
        from oracle import oracle as orc
        rng = np.random.default_rng(123)                      # same genome on every rank
        genome = "".join(rng.choice(list("ACGT"), size=3000))
        reads = [genome[s:s + 60] for s in rng.integers(0, len(genome) - 60, size=400)]
        mine = reads[rank::world]
        kmers = orc.extract(mine, K)                          # this rank's canonical k-mer instances
        dest = np.array([owner_of(int(x), K, M, world) for x in kmers], dtype=np.int64)
        lbin = (kmers % np.uint64(P_LOCAL)).astype(np.int64)  # any deterministic local bin
        P = world * P_LOCAL
        key = dest * P_LOCAL + lbin
        order = np.argsort(key, kind="stable")
        send = kmers[order]                                   # destination-major, bin-minor: what K1's scatter produces
        cnt = np.bincount(key, minlength=P).astype(np.uint64)
        allc = [torch.zeros(P, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, torch.from_numpy(cnt.astype(np.int64)))
        all_cnt = np.stack([a.numpy().astype(np.uint64) for a in allc])
        send_cnt, recv_cnt, run_off, run_len = plan_shard_exchange(world, rank, P_LOCAL, all_cnt)
        assert int(send_cnt.sum()) == len(kmers)
        # transport: every rank publishes its send buffer, receivers slice their segment (stands in for ncclSend/ncclRecv)
        bufs = [None] * world
        dist.all_gather_object(bufs, send)
        recv = []
        for s in range(world):
            seg_start = int(all_cnt[s, : rank * P_LOCAL].sum())
            recv.append(bufs[s][seg_start: seg_start + int(recv_cnt[s])])
        recv = np.concatenate(recv) if recv else np.zeros(0, np.uint64)
        assert len(recv) == int(recv_cnt.sum())
        # every received k-mer belongs here, and the runs of bin b hold exactly the k-mers of local bin b
        assert all(owner_of(int(x), K, M, world) == rank for x in recv)
        seen = 0
        for b in range(P_LOCAL):
            for s in range(world):
                run = recv[int(run_off[b, s]): int(run_off[b, s]) + int(run_len[b, s])]
                assert np.all(run % np.uint64(P_LOCAL) == b)
                seen += len(run)
        assert seen == len(recv)
        # global check on rank 0: the shards partition the multiset of all instances and no k-mer is split between ranks
        allrecv = [None] * world
        dist.all_gather_object(allrecv, recv)
        if rank == 0:
            total = np.sort(np.concatenate(allrecv))
            ref = np.sort(orc.extract(reads, K))
            assert np.array_equal(total, ref)
            owners = {}
            for r, arr in enumerate(allrecv):
                for x in np.unique(arr):
                    assert owners.setdefault(int(x), r) == r
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"FAIL: {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_exchange_plan_two_process_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
