# dev helper: C2-scale run on the GPU with stage timings (not part of the product)
import sys, time, json
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bcalm_b200.api import Cdbg, MEM_DEVICE
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
K = int(os.environ.get('DEV_K', '31')); L = int(os.environ.get('DEV_L', '100'))
G = int(5_000_000 * scale); n_reads = int(G * 30 / L)
c = Cdbg(k=K, abundance_min=2)
ptr, n = c.synth_reads(G, L, n_reads, 10000, 42)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for it in range(iters):
    t0 = time.perf_counter()
    v, s = c.build_raw(ptr, n, MEM_DEVICE, fetch=True)
    dt = time.perf_counter() - t0
    st = c.stats()
    print(f"iter {it}: wall {dt*1e3:.2f} ms  {n_reads*L/dt/1e9:.2f} Gbases/s")
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}))
