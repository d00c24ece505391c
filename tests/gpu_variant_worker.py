"""Worker of tests/test_gpu_variants.py: one process = one CUDA context = one configuration.

    python tests/gpu_variant_worker.py variant  NAME=VALUE ...      build golden cases + a 7.5 Mbase synthetic set under the
                                                                    environment given, compare with the reference's golden output /
                                                                    the default configuration
    python tests/gpu_variant_worker.py features                     the entry points added late in round 1 (counted k-mers in, FASTA /
                                                                    FASTQ text in, even-k dense graphs)
Prints one line per check and exits non-zero on the first mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bcalm_b200 import parity  # noqa: E402
import bcalm_b200.api as _api  # noqa: E402
from bcalm_b200.api import Cdbg, MEM_DEVICE  # noqa: E402

if os.environ.get("CDBG_GPU_TEST_LIB"):      # scripts/dry_run_gpu_tests.py: the tests' own logic against the CPU emulation
    _api.LIB_PATH = os.environ["CDBG_GPU_TEST_LIB"]

with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
    GOLDEN = {c["name"]: c for c in json.load(f)}
CASES = ["tiny_read_k21", "circular_test1", "lowcx_8_ACGT_k8_a1", "palindrome_k_even", "pow2_cycle_L256_k21", "reads_3_k31_a2", "reads_5_k25_a1",
         "even_k_dense_k6"]


def check_golden(name):
    c = GOLDEN[name]
    with Cdbg(k=c["k"], abundance_min=c["abundance_min"]) as x:
        g = x.build(c["reads"])
    d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
    print(f"{name}: {'OK' if d == '' else 'MISMATCH ' + d[:300]}", flush=True)
    return d == ""


def synthetic_summary(env):
    """250 kbp genome, 30x, 100 bp reads: sorted solid k-mers + graph summary, with and without `env`."""
    out = []
    for e in ({}, env):
        saved = {k: os.environ.pop(k, None) for k in env}
        os.environ.update(e)
        with Cdbg(k=31, abundance_min=2) as c:
            ptr, n = c.synth_reads(250_000, 100, 75_000, 10000, 7)
            c.build_raw(ptr, n, MEM_DEVICE, fetch=False)
            km, ct = c.solid_kmers()
            g = c.unitigs()
            o = np.argsort(km)
            out.append((km[o], ct[o], g.n_unitigs, len(g.links), int(g.kc.sum()), c.stats()["ms_count"] + c.stats()["ms_compact"]))
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    a, b = out
    same = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:5] == b[2:5]
    print(f"synthetic 7.5 Mbases: {'SAME' if same else 'DIFFERENT'} as the default configuration; stage ms default {a[5]:.3f} variant {b[5]:.3f}", flush=True)
    return same


def main():
    mode = sys.argv[1]
    ok = True
    if mode == "variant":
        env = dict(a.split("=", 1) for a in sys.argv[2:])
        os.environ.update(env)
        for name in CASES:
            ok = check_golden(name) and ok
        ok = synthetic_summary(env) and ok
    else:
        for name in ("even_k_dense_k4", "even_k_dense_k6", "even_k_dense_AT_k8", "even_k_dense_ACG_k6", "even_k_reads_k20_a2"):
            ok = check_golden(name) and ok
        c = GOLDEN["reads_3_k31_a2"]
        with Cdbg(k=c["k"], abundance_min=c["abundance_min"]) as a:
            a.push_reads(c["reads"]); a.count(); km, ct = a.solid_kmers()
        with Cdbg(k=c["k"], abundance_min=1) as b:
            b.set_solid_kmers(km, ct); b.compact(); b.link(); g = b.unitigs()
        d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
        print("cdbg_set_solid_kmers:", "OK" if d == "" else "MISMATCH " + d[:300], flush=True)
        ok = ok and d == ""
        fa = "".join(f">r{i} ACGTACGTACGTACGTACGTACGTACGTACGTACGT\n{r[:33]}\r\n{r[33:]}\n" for i, r in enumerate(c["reads"]))
        fq = "".join(f"@r{i}\n{r}\n+\n{'@' * len(r)}\n" for i, r in enumerate(c["reads"]))
        for label, text in (("FASTA", fa), ("FASTQ", fq)):
            with Cdbg(k=c["k"], abundance_min=c["abundance_min"]) as x:
                n = x.push_sequence_text(text); x.count(); x.compact(); x.link(); g = x.unitigs()
            d = parity.diff(parity.graph_to_records(g), parity.parse_unitigs_fa(c["unitigs_fa"]), c["k"])
            good = d == "" and n == len(c["reads"])
            print(f"cdbg_push_sequence_text {label}:", "OK" if good else f"MISMATCH records={n} " + d[:300], flush=True)
            ok = ok and good
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
