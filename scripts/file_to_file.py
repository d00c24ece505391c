#!/usr/bin/env python
"""file -> file timing of the bcalm-compatible CLI next to the reference binary (same FASTA, same flags), C2 workload."""
import os, subprocess, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bcalm_b200 import parity
from bcalm_b200.api import Cdbg
from oracle import oracle as orc
sys.path.insert(0, ROOT)
import bench

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
g, n_reads = bench.workload(scale)
with Cdbg(k=31, abundance_min=2) as c:
    ptr, n = c.synth_reads(g, bench.READ_LEN, n_reads, bench.ERR_PPM, bench.SEED)
    reads = c.copy_to_host(ptr, n)
out = {}
with tempfile.TemporaryDirectory() as d:
    fa = os.path.join(d, "c2.fa")
    bench.write_fasta_from_bytes(fa, reads, bench.READ_LEN)
    cli = os.path.join(ROOT, "bcalm_b200", "bcalm")
    for it in range(3):
        t0 = time.perf_counter()
        p = subprocess.run([cli, "-in", fa, "-kmer-size", "31", "-abundance-min", "2", "-out", os.path.join(d, "ours")], capture_output=True, text=True, cwd=d)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stdout
    out["ours_s"] = dt
    out["ours_log"] = [l for l in p.stdout.splitlines() if l.startswith("wall s") or l.startswith("device ms")]
    if orc.have_ref():
        text, dtr, _ = orc.run_ref(fa, 31, 2, nb_cores=os.cpu_count(), extra=["-verbose", "0"])
        out["reference_s"] = dtr
        out["cores"] = os.cpu_count()
        ours = open(os.path.join(d, "ours.unitigs.fa")).read()
        diff = parity.diff(parity.parse_unitigs_fa(ours), parity.parse_unitigs_fa(text), 31)
        out["parity_vs_reference_binary"] = "equal" if diff == "" else diff[:300]
out["mbases"] = n_reads * bench.READ_LEN / 1e6
out["ours_gbases_s"] = out["mbases"] / 1e3 / out["ours_s"]
if "reference_s" in out:
    out["reference_gbases_s"] = out["mbases"] / 1e3 / out["reference_s"]
print(json.dumps(out))
