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
    ("default: K1 single, K2 smem, K5 l2, atomic ranks", {}),
    ("k2c flat insert loop", {"CDBG_K2C": "flat"}),
    ("k2c mul32 slot hash", {"CDBG_K2C_HASH": "mul32"}),
    ("k2c mul32 slot hash + flat", {"CDBG_K2C_HASH": "mul32", "CDBG_K2C": "flat"}),
    ("k2c flat, fill 1.0", {"CDBG_K2C": "flat", "CDBG_K2S_FILL": "1.0"}),
    ("k2c flat, fill 2.0", {"CDBG_K2C": "flat", "CDBG_K2S_FILL": "2.0"}),
    ("K2 fine partition by regions (no hist pass)", {"CDBG_K2P": "regions"}),
    ("K2 regions + k2c flat + K1 512", {"CDBG_K2P": "regions", "CDBG_K2C": "flat", "CDBG_K1_THREADS": "512"}),
    ("K1 512 threads", {"CDBG_K1_THREADS": "512"}),
    ("K1 512 threads + k2c flat", {"CDBG_K1_THREADS": "512", "CDBG_K2C": "flat"}),
    ("ranks by match.any", {"CDBG_RANK": "match"}),
    ("K5 smem 2048/1.0", {"CDBG_K5": "smem"}),
    ("K5 smem 2048/0.7", {"CDBG_K5": "smem", "CDBG_K5S_FILL": "0.7"}),
    ("K5 smem 2048/1.4", {"CDBG_K5": "smem", "CDBG_K5S_FILL": "1.4"}),
    ("K5 smem 4096/1.0", {"CDBG_K5": "smem", "CDBG_K5S_SLOTS": "4096"}),
    ("K5 smem 2048/1.0 P1=32", {"CDBG_K5": "smem", "CDBG_K5S_P1": "32"}),
    ("K5 smem + match ranks", {"CDBG_K5": "smem", "CDBG_RANK": "match"}),
    ("K6 over lists", {"CDBG_K6": "list"}),
    ("K5 smem + K6 lists", {"CDBG_K5": "smem", "CDBG_K6": "list"}),
    ("K5 smem + K6 lists + match ranks", {"CDBG_K5": "smem", "CDBG_K6": "list", "CDBG_RANK": "match"}),
    ("round-start: K1 twopass, K2 l2, K5 l2", {"CDBG_K1": "twopass", "CDBG_K2": "l2"}),
]
E2E_CONFIGS = [
    ("e2e default (zero-copy, single, K2 smem, K5 l2)", {}),
    ("e2e + K1 512 threads + k2c flat", {"CDBG_K1_THREADS": "512", "CDBG_K2C": "flat"}),
    ("e2e + K5 smem", {"CDBG_K5": "smem"}),
    ("e2e + K5 smem + K6 lists", {"CDBG_K5": "smem", "CDBG_K6": "list"}),
    ("e2e + K5 smem + K6 lists + match ranks", {"CDBG_K5": "smem", "CDBG_K6": "list", "CDBG_RANK": "match"}),
    ("e2e round-start (copy, twopass, l2)", {"CDBG_K1": "twopass", "CDBG_K2": "l2", "CDBG_ZEROCOPY": "0"}),
]
TUNABLES = TUNABLES + ("CDBG_ZEROCOPY", "CDBG_K1_SLACK", "CDBG_K5", "CDBG_K5S_FILL", "CDBG_K5S_SLOTS", "CDBG_K5S_P1", "CDBG_RANK", "CDBG_K6", "CDBG_K2C", "CDBG_K1_THREADS", "CDBG_K2P", "CDBG_K2C_HASH")


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
            f"launches {st['kernel_launches']} | K2 parts hist {st['ms_k2_hist']:.3f} scatter {st['ms_k2_scatter']:.3f} table {st['ms_k2_table']:.3f} "
            f"splits {st['n_k2_splits']}/{st['n_k5_splits']} k1fb {st['k1_fallback']} rounds {st['rank_rounds']} | S,D,U,nt,links,KC = {summary}")
    except Exception as e:  # keep going: one bad configuration must not cost the whole call
        say(f"{name:50s} FAILED: {e}")
        try:
            c.close()
        except Exception:
            pass
        c = Cdbg(k=K, abundance_min=2)
        ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)

# ---- end to end through cdbg_build with the reads in pinned host memory (H2D inside the timed region, result fetched)
import ctypes as C  # noqa: E402

lib = api.load_library()
hptr = lib.cdbg_host_alloc(n)
assert hptr, "pinned allocation failed"
c._ck(lib.cdbg_copy_to_host(c.h, C.c_void_p(hptr), C.c_void_p(ptr), n))
for name, env in E2E_CONFIGS:
    for k_ in TUNABLES:
        os.environ.pop(k_, None)
    os.environ.update(env)
    try:
        ts = []
        for it in range(7):
            t0 = time.perf_counter()
            v, s = c.build_raw(hptr, n, 0, fetch=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        st = c.stats()
        ts = sorted(ts[2:])
        ok = (int(st["n_solid"]), int(v.n_unitigs), int(v.n_links)) == (base[2][0], base[2][2], base[2][4]) if base else None
        say(f"{name:50s} wall min {ts[0]:.3f} median {ts[len(ts) // 2]:.3f} ms -> {n_reads * L / ts[len(ts) // 2] / 1e6:.2f} Gbases/s | zero_copy {st['zero_copy']} "
            f"push {st['ms_push']:.3f} count {st['ms_count']:.3f} (K1 {st['ms_k1_hist']:.3f}+{st['ms_k1_scatter']:.3f}) compact {st['ms_compact']:.3f} link {st['ms_link']:.3f} "
            f"fetch {st['ms_fetch']:.3f} | same counts: {ok}")
    except Exception as e:
        say(f"{name:50s} FAILED: {e}")
lib.cdbg_host_free(C.c_void_p(hptr))
say(f"# done in {time.time() - T0:.1f} s")
