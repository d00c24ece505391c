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
    "k2c_flat_insert_loop": ["CDBG_K2C=flat"],
    "k2c_mul32_slot_hash": ["CDBG_K2C_HASH=mul32"],
    "k1_512_threads": ["CDBG_K1_THREADS=512"],
    "k1_512_threads_and_k2c_flat": ["CDBG_K1_THREADS=512", "CDBG_K2C=flat"],
    "k2p_regions": ["CDBG_K2P=regions"],
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
