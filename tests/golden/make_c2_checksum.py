"""Writes tests/golden/c2_checksum.json: the checksum (bcalm_b200/checksum.py) of the C ORACLE's graph of bench.py's C2 reads at the
scales bench.py and the tests use.  bench.py compares the GPU graph of the same reads with it (its N = 1 "parity" key) without
touching oracle/; tests/test_gpu_full_scale.py compares the GPU graph with the oracle itself.

    python tests/golden/make_c2_checksum.py            (about two minutes on one core: the oracle runs full C2)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from bcalm_b200.checksum import graph_checksum  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out = {}
for scale in (0.002, 0.01, 0.2, 1.0):
    g, n_reads = bench.workload(scale)
    reads = bench.synth_reads_numpy(g, bench.READ_LEN, 0, n_reads, bench.ERR_PPM, bench.SEED)
    o = orc.run(reads, bench.K, bench.AMIN)
    cs = graph_checksum(o)
    cs["n_solid"] = int(o.stats["n_solid"])
    out[f"scale_{scale:g}"] = cs
    print(scale, cs, flush=True)
with open(os.path.join(ROOT, "tests", "golden", "c2_checksum.json"), "w") as f:
    json.dump(out, f, indent=1)
