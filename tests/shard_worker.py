"""Worker for the sharded (multi-GPU) parity tests; launched by torchrun, one process per GPU.

    torchrun --nproc-per-node N tests/shard_worker.py <case names | c2:<scale>>

Every rank takes a round-robin share of the reads, all ranks build collectively, rank 0 merges the shards and compares
with the oracle (and with the reference's golden output when the case has one)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bcalm_b200 import parity  # noqa: E402
from bcalm_b200.api import Cdbg, merge_shards  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        golden = {c["name"]: c for c in json.load(f)}
    failures = []
    # cases that share (k, abundance-min, flags) share one context (one NCCL communicator): sort so that they are consecutive
    def _key(name):
        ab = name.startswith("ab:")
        n = name[3:] if ab else name
        if n.startswith("c2:"):
            return (31, 2, ab, name)
        return (golden[n]["k"], golden[n]["abundance_min"], ab, name)
    cur_key, cd = None, None
    for name in sorted(sys.argv[1:], key=_key):
        ctx_key = _key(name)[:3]
        want_ab = name.startswith("ab:")             # ab:<case> = the same case with -all-abundance-counts (per-k-mer lists checked too)
        if want_ab:
            name = name[3:]
        if name.startswith("c2:"):
            scale = float(name[3:])
            G = int(5_000_000 * scale); L = 100; n_reads = G * 30 // L
            k, amin = 31, 2
            with Cdbg(k=k, abundance_min=amin, device=local) as tmp:
                ptr, n = tmp.synth_reads(G, L, n_reads, 10000, 42)
                allreads = tmp.copy_to_host(ptr, n).reshape(n_reads, L + 1)
            mine = np.ascontiguousarray(allreads[rank::world]).reshape(-1)
            full = allreads.reshape(-1)
            ref_text = None
        else:
            c = golden[name]
            k, amin = c["k"], c["abundance_min"]
            reads = c["reads"]
            mine = np.frombuffer(("\n".join(reads[rank::world]) + "\n").encode(), dtype=np.uint8) if reads[rank::world] else np.zeros(0, np.uint8)
            full = np.frombuffer(("\n".join(reads) + "\n").encode(), dtype=np.uint8)
            ref_text = c["unitigs_fa"]
        if ctx_key != cur_key:
            if cd is not None:
                cd.close()
            cd = Cdbg(k=k, abundance_min=amin, device=local, rank=rank, world_size=world, all_abundance_counts=want_ab)
            cd.shard_init_torch()
            cur_key = ctx_key
        g = cd.build(mine)
        gathered = [None] * world
        dist.gather_object(g, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            from oracle import oracle as orc
            m = merge_shards(gathered)
            if k <= 31:
                o = orc.run(full, k, amin)
                d = parity.diff(parity.graph_to_records(m), parity.graph_to_records(o), k)
                n_solid_expected = o.stats["n_solid"]
            else:                                   # the C oracle is k <= 31: multi-word cases are checked against the reference's golden output
                assert ref_text is not None, "k > 31 needs a golden case"
                o, d = None, ""
                n_solid_expected = sum(r["ln"] - k + 1 for r in parity.parse_unitigs_fa(ref_text))
            if not d and ref_text:
                d = parity.diff(parity.graph_to_records(m), parity.parse_unitigs_fa(ref_text), k)
            if not d and want_ab and o is not None:
                d = "no abundance lists in the merged shards" if m.abund is None else \
                    parity.diff_abundance_lists(parity.graph_to_records(m, with_ab=True), parity.graph_to_records(o, with_ab=True))
            solid = sum(s["n_solid"] for s in m.stats["shards"])
            ok = (d == "") and solid == n_solid_expected
            print(f"[shard x{world}] {'ab:' if want_ab else ''}{name}: {'OK' if ok else 'FAIL'} unitigs={m.n_unitigs} links={len(m.links)} solid={solid}/{n_solid_expected} "
                  f"per-rank unitigs={[x.n_unitigs for x in gathered]} pieces={[x.stats['n_unitigs'] for x in gathered]}", flush=True)
            if not ok:
                print(d[:2000], flush=True)
                failures.append(name)
        dist.barrier()
    if cd is not None:
        cd.close()
    if rank == 0:
        print("SHARD_RESULT", "FAIL " + ",".join(failures) if failures else "PASS", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
