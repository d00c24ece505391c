# dev helper (not part of the product): A/B of stage 1 on super-k-mers against the k-mer-instance path, C2 and a C3-like slice.
#   python scripts/dev_sk_ab.py [genome_bp] [read_len] > gpurun_out/sk_ab.txt
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bcalm_b200 import api  # noqa: E402
from bcalm_b200.api import Cdbg, MEM_DEVICE  # noqa: E402

if os.environ.get("AB_LIB"):
    api.LIB_PATH = os.environ["AB_LIB"]
G = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
K = 31
n_reads = G * 30 // L
T0 = time.time()
BUDGET_S = float(os.environ.get("AB_BUDGET_S", "90"))
TUN = ("CDBG_S1", "CDBG_SK_M", "CDBG_SK_SLOTS", "CDBG_SK_FILL", "CDBG_SK_P1", "CDBG_K2C_HASH", "CDBG_K1_THREADS", "CDBG_SK_DENSITY", "CDBG_K6", "CDBG_ZEROCOPY", "CDBG_SK_KC", "CDBG_SK2C", "CDBG_SK_RS", "CDBG_MMER_RANK", "CDBG_K5S_FILL", "CDBG_K5S_SLOTS", "CDBG_SK_INSERT", "CDBG_SK_THREADS", "CDBG_SK_WCAP")
CONFIGS = [
    ("k-mer instances (round-1 default)", {"CDBG_S1": "kmer"}),
    ("super-k-mers default", {}),
    ("sk two-phase insert", {"CDBG_SK_INSERT": "two"}),
    ("sk two-phase insert mix64", {"CDBG_SK_INSERT": "two", "CDBG_K2C_HASH": "mix64"}),
    ("sk two-phase insert fill 2.0", {"CDBG_SK_INSERT": "two", "CDBG_SK_FILL": "2.0"}),
    ("sk two-phase insert fill 1.0", {"CDBG_SK_INSERT": "two", "CDBG_SK_FILL": "1.0"}),
    ("sk small staging (4 CTAs/SM)", {"CDBG_SK_WCAP": "128"}),
    ("sk small staging fill 1.25", {"CDBG_SK_WCAP": "128", "CDBG_SK_FILL": "1.25"}),
    ("sk 512 threads per table", {"CDBG_SK_THREADS": "512"}),
    ("sk 512 threads fill 2.0", {"CDBG_SK_THREADS": "512", "CDBG_SK_FILL": "2.0"}),
    ("sk 512 threads loop insert", {"CDBG_SK_THREADS": "512", "CDBG_SK_INSERT": "loop"}),
    ("sk K6 full (one level)", {"CDBG_K6": "full"}),
    ("super-k-mers direct count", {"CDBG_SK2C": "direct"}),
    ("super-k-mers dedup RS 1024", {"CDBG_SK2C": "dedup"}),
    ("super-k-mers dedup RS 2048", {"CDBG_SK2C": "dedup", "CDBG_SK_RS": "2048"}),
    ("super-k-mers dedup RS 512", {"CDBG_SK2C": "dedup", "CDBG_SK_RS": "512"}),
    ("super-k-mers dedup mul32", {"CDBG_SK2C": "dedup", "CDBG_K2C_HASH": "mul32"}),
    ("super-k-mers dedup fill 2.0", {"CDBG_SK2C": "dedup", "CDBG_SK_FILL": "2.0"}),
    ("super-k-mers dedup fill 3.0 RS 2048", {"CDBG_SK2C": "dedup", "CDBG_SK_FILL": "3.0", "CDBG_SK_RS": "2048"}),
    ("super-k-mers dedup 8192 fill 2 RS 2048", {"CDBG_SK2C": "dedup", "CDBG_SK_SLOTS": "8192", "CDBG_SK_FILL": "2.0", "CDBG_SK_RS": "2048"}),
    ("sk + K6 two-level", {"CDBG_K6": "local"}),
    ("sk dedup + K6 two-level", {"CDBG_SK2C": "dedup", "CDBG_K6": "local"}),
    ("sk dedup fill 2 + K6 two-level", {"CDBG_SK2C": "dedup", "CDBG_K6": "local", "CDBG_SK_FILL": "2.0"}),
    ("super-k-mers runtime k", {"CDBG_SK_KC": "0"}),
    ("super-k-mers mul32", {"CDBG_K2C_HASH": "mul32"}),
    ("super-k-mers fill 1.0", {"CDBG_SK_FILL": "1.0"}),
    ("super-k-mers fill 2.0", {"CDBG_SK_FILL": "2.0"}),
    ("super-k-mers fill 2.5", {"CDBG_SK_FILL": "2.5"}),
    ("super-k-mers slots 8192", {"CDBG_SK_SLOTS": "8192"}),
    ("super-k-mers slots 8192 fill 2.0", {"CDBG_SK_SLOTS": "8192", "CDBG_SK_FILL": "2.0"}),
    ("super-k-mers slots 8192 mul32", {"CDBG_SK_SLOTS": "8192", "CDBG_K2C_HASH": "mul32"}),
    ("super-k-mers m=13", {"CDBG_SK_M": "13"}),
    ("super-k-mers m=15", {"CDBG_SK_M": "15"}),
    ("super-k-mers m=16", {"CDBG_SK_M": "16"}),
]
if os.environ.get("AB_ONLY"):
    want = os.environ["AB_ONLY"].split(",")
    CONFIGS = [c for c in CONFIGS if any(w in c[0] for w in want)]


def say(*a):
    print(*a, flush=True)


c = Cdbg(k=K, abundance_min=2)
ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
say(f"# genome {G} bp, {n_reads} reads x {L} bp, {n} bytes resident")
base = None
for name, env in CONFIGS:
    if time.time() - T0 > BUDGET_S:
        say(f"# budget used up before '{name}'")
        break
    for k_ in TUN:
        os.environ.pop(k_, None)
    os.environ.update(env)
    try:
        best = None
        for it in range(4):
            v, s = c.build_raw(ptr, n, MEM_DEVICE, fetch=False)
            st = c.stats()
            tot = st["ms_count"] + st["ms_compact"] + st["ms_link"]
            if it and (best is None or tot < best[0]):
                best = (tot, st)
        tot, st = best
        km, ct = c.solid_kmers()
        order = np.argsort(km, kind="stable")
        km, ct = km[order], ct[order]
        g = c.unitigs(copy=False)
        summary = (int(st["n_solid"]), int(st["n_distinct"]), int(g.n_unitigs), int(len(g.seq)), int(len(g.links)), int(g.kc.sum()))
        if base is None:
            base = (km, ct, summary)
            same = "baseline"
        else:
            same = "SAME" if (summary == base[2] and np.array_equal(km, base[0]) and np.array_equal(ct, base[1])) else "DIFFERENT"
        say(f"{name:40s} {same:9s} total {tot:7.3f} ms count {st['ms_count']:7.3f} | K1 {st['ms_k1_hist']:.3f}+{st['ms_k1_scatter']:.3f} K2 {st['ms_k2']:.3f} "
            f"(hist {st['ms_k2_hist']:.3f} scatter {st['ms_k2_scatter']:.3f} table {st['ms_k2_table']:.3f}) K3 {st['ms_k3']:.3f} K5 {st['ms_k5']:.3f} "
            f"K6 {st['ms_k6_rank']:.3f}+{st['ms_k6_spell']:.3f} link {st['ms_link']:.3f} | bins {st['n_count_bins']} m {st['sk_m']} sk {st['n_superkmers']} "
            f"I {st['n_instances']} splits {st['n_k2_splits']} k1fb {st['k1_fallback']} rounds {st['rank_rounds']} launches {st['kernel_launches']} | {summary}")
    except Exception as e:
        say(f"{name:40s} FAILED: {e}")
        try:
            c.close()
        except Exception:
            pass
        c = Cdbg(k=K, abundance_min=2)
        ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
say(f"# done in {time.time() - T0:.1f} s")
