"""GPU checks of the opt-in kernel variants and of the late entry points (counted k-mers in, FASTA/FASTQ text in).
Each check runs in its OWN process, so a misbehaving variant cannot damage the CUDA context of the suite.  All of them passed on
hardware (GPUTEST_r01: xpassed) and gate like any other test since round 2."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "gpu_variant_worker.py")

VARIANTS = {
    # stage 1 on k-mer instances (the round-1 path) and its variants
    "s1_kmer_instances": ["CDBG_S1=kmer"],
    "k2c_flat_insert_loop": ["CDBG_S1=kmer", "CDBG_K2C=flat"],
    "k2c_mix64_slot_hash": ["CDBG_S1=kmer", "CDBG_K2C_HASH=mix64"],
    "k1_512_threads": ["CDBG_S1=kmer", "CDBG_K1_THREADS=512"],
    "k1_512_threads_and_k2c_flat": ["CDBG_S1=kmer", "CDBG_K1_THREADS=512", "CDBG_K2C=flat"],
    "k2p_regions": ["CDBG_S1=kmer", "CDBG_K2P=regions"],
    # variants of the super-k-mer path and of the ranking (round 2)
    "sk_dedup_count": ["CDBG_SK2C=dedup"],
    "sk_loop_insert_mix64": ["CDBG_SK_INSERT=loop", "CDBG_K2C_HASH=mix64"],
    "sk_three_waves": ["CDBG_SK_WAVES=3"],
    "sk_512_threads_per_table": ["CDBG_SK_THREADS=512"],
    "k6_one_level": ["CDBG_K6=full"],
    "k6_one_level_lists": ["CDBG_K6=list"],
    "k5_l2_tables": ["CDBG_K5=l2"],
}
WORKER_TIMEOUT_S = 90      # a healthy worker needs ~10 s; a variant that hangs must not hold up the suite for long


def run_worker(*args):
    p = subprocess.run([sys.executable, WORKER, *args], capture_output=True, text=True, timeout=WORKER_TIMEOUT_S)
    print(p.stdout[-4000:])
    print(p.stderr[-2000:], file=sys.stderr)
    return p.returncode


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variant_gives_identical_output_on_the_gpu(variant):
    assert run_worker("variant", *VARIANTS[variant]) == 0


def test_late_entry_points_on_the_gpu():
    assert run_worker("features") == 0
