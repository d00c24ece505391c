# dev helper (not part of the product): A/B of the K2 variants and a few tunables on the C2 workload, one short GPU call.
#   python scripts/dev_k2_ab.py [scale] > gpurun_out/k2_ab.txt
# For every configuration (environment variables read by the library at each cdbg_count/cdbg_compact): 1 warm-up + 3 timed
# builds with the reads resident in HBM, device time per stage from the library's CUDA events, and the solid (k-mer, count)
# set + the graph summary compared with the default configuration's.  No torch import: the call has to fit in a minute.
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from bcalm_b200 import api  # noqa: E402
from bcalm_b200.api import Cdbg, MEM_DEVICE  # noqa: E402

if os.environ.get("AB_LIB"):   # dry run of this script against the CPU emulation build (tests/emu)
    api.LIB_PATH = os.environ["AB_LIB"]

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
T0 = time.time()
BUDGET_S = float(os.environ.get("AB_BUDGET_S", "75"))
K, L = 31, 100
G = int(5_000_000 * scale)
n_reads = int(G * 30 / L)
TUNABLES = ("CDBG_K2", "CDBG_K2S_SLOTS", "CDBG_K2S_FILL", "CDBG_K2S_P1", "CDBG_JOIN_TARGET", "CDBG_BIN_TARGET", "CDBG_K1")

CONFIGS = [
    ("default (L2 tables)", {}),
    ("smem 8192 fill 0.7", {"CDBG_K2": "smem"}),
    ("smem 8192 fill 1.5", {"CDBG_K2": "smem", "CDBG_K2S_FILL": "1.5"}),
    ("smem 8192 fill 2.2", {"CDBG_K2": "smem", "CDBG_K2S_FILL": "2.2"}),
    ("smem 4096 fill 0.7", {"CDBG_K2": "smem", "CDBG_K2S_SLOTS": "4096"}),
    ("smem 4096 fill 1.5", {"CDBG_K2": "smem", "CDBG_K2S_SLOTS": "4096", "CDBG_K2S_FILL": "1.5"}),
    ("smem 16384 fill 0.7", {"CDBG_K2": "smem", "CDBG_K2S_SLOTS": "16384"}),
    ("smem 16384 fill 1.5", {"CDBG_K2": "smem", "CDBG_K2S_SLOTS": "16384", "CDBG_K2S_FILL": "1.5"}),
    ("smem 16384 fill 2.2", {"CDBG_K2": "smem", "CDBG_K2S_SLOTS": "16384", "CDBG_K2S_FILL": "2.2"}),
    ("smem 8192 fill 1.5 P1=64", {"CDBG_K2": "smem", "CDBG_K2S_FILL": "1.5", "CDBG_K2S_P1": "64"}),
    ("smem 8192 fill 1.5 P1=256", {"CDBG_K2": "smem", "CDBG_K2S_FILL": "1.5", "CDBG_K2S_P1": "256"}),
    ("smem 8192 fill 6 (forces the overflow fallback)", {"CDBG_K2": "smem", "CDBG_K2S_FILL": "6"}),
    ("join target 1M", {"CDBG_JOIN_TARGET": "1000000"}),
    ("join target 2M", {"CDBG_JOIN_TARGET": "2000000"}),
    ("join target 4M", {"CDBG_JOIN_TARGET": "4000000"}),
    ("join target 250k", {"CDBG_JOIN_TARGET": "250000"}),
]


def say(*a):
    print(*a, flush=True)


c = Cdbg(k=K, abundance_min=2)
ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
say(f"# C2 x {scale}: {n_reads} reads x {L} bp, {n} bytes resident; expected at scale 1: S=5598931 U=85395 links=155028")
base = None
for name, env in CONFIGS:
    if time.time() - T0 > BUDGET_S:
        say(f"# budget of {BUDGET_S:.0f} s used up before '{name}'")
        break
    for k_ in TUNABLES:
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
        say(f"{name:50s} {same:9s} total {tot:6.3f} ms | K1 {st['ms_k1_hist']:.3f}+{st['ms_k1_scatter']:.3f} K2 {st['ms_k2']:.3f} K3 {st['ms_k3']:.3f} "
            f"K5 {st['ms_k5']:.3f} K6 {st['ms_k6_rank']:.3f}+{st['ms_k6_spell']:.3f} link {st['ms_link']:.3f} | bins {st['n_count_bins']}/{st['n_join_bins']} "
            f"launches {st['kernel_launches']} | S,D,U,nt,links,KC = {summary}")
    except Exception as e:  # keep going: one bad configuration must not cost the whole call
        say(f"{name:50s} FAILED: {e}")
        try:
            c.close()
        except Exception:
            pass
        c = Cdbg(k=K, abundance_min=2)
        ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
say(f"# done in {time.time() - T0:.1f} s")
