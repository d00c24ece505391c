# dev helper: per-step e2e timing (pinned host reads -> cdbg_build -> fetched unitigs) with the library's stage timers
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bcalm_b200.api import Cdbg, MEM_HOST, MEM_DEVICE
G, L = 5_000_000, 100
n_reads = G * 30 // L
for env in ({}, {"CDBG_K6": "full"}, {"CDBG_K6": "full", "CDBG_K2C_HASH": "mix64"}, {"CDBG_ZEROCOPY": "0"}):
    for k_ in ("CDBG_K6", "CDBG_K2C_HASH", "CDBG_ZEROCOPY"):
        os.environ.pop(k_, None)
    os.environ.update(env)
    c = Cdbg(k=31, abundance_min=2)
    ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy(); hnp[:] = c.copy_to_host(ptr, n)
    ts = []
    for it in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, s = c.build_raw(hnp.ctypes.data, n, MEM_HOST, fetch=True)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    st = s.as_dict()
    print(env, "e2e ms per step:", " ".join(f"{t:.2f}" for t in ts))
    print("   ", {k_: round(v_, 3) for k_, v_ in st.items() if k_.startswith("ms_")}, "zero_copy", st["zero_copy"], "launches", st["kernel_launches"], flush=True)
    c.close()
