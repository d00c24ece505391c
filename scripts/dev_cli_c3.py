# dev helper: the CLI file -> file on a C3-sized FASTA (250 Mbp genome, 50 M x 150 bp reads, 7.8 GB of text), next to the C2 figure of bench.py
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from bcalm_b200.api import Cdbg
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
G, L = int(250_000_000 * scale), 150
n_reads = G * 30 // L
d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
fa = os.path.join(d, "c3.fa")
t0 = time.time()
with Cdbg(k=31, abundance_min=2) as c:
    with open(fa, "wb") as f:
        step = 5_000_000
        for first in range(0, n_reads, step):
            n = min(step, n_reads - first)
            ptr, nb = c.synth_reads(G, L, n, 10000, 43, first_read=first)
            a = c.copy_to_host(ptr, nb).reshape(n, L + 1)
            out = np.empty((n, L + 4), np.uint8)
            out[:, 0] = ord(">"); out[:, 1] = ord("r"); out[:, 2] = ord("\n"); out[:, 3:] = a
            out.tofile(f)
print(f"# wrote {os.path.getsize(fa) / 1e9:.2f} GB FASTA in {time.time() - t0:.1f} s", flush=True)
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bcalm_b200", "bcalm")
for mode in ("1", "0"):
    best = None
    for it in range(2):
        t0 = time.perf_counter()
        p = subprocess.run([cli, "-in", fa, "-kmer-size", "31", "-abundance-min", "2", "-out", os.path.join(d, "out"), "-gpu-parse", mode], capture_output=True, text=True, cwd=d)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            print("CLI failed:", p.stdout[-400:]); break
        line = [l for l in p.stdout.splitlines() if l.startswith("wall s")]
        if best is None or dt < best[0]:
            best = (dt, line[0] if line else "")
    if best:
        print(f"-gpu-parse {mode}: {n_reads * L / best[0] / 1e9:.3f} Gbases/s file -> file, {best[0]:.2f} s | {best[1]}", flush=True)
    os.remove(os.path.join(d, "out.unitigs.fa"))
os.remove(fa)
