"""GPU checks of code that could only be verified under the CPU emulator when it was written (the round's GPU budget was spent):
the opt-in kernel variants and the entry points added late (counted k-mers in, FASTA/FASTQ text in, the even-k
join fix's golden cases).  Each check runs in its OWN process, so a misbehaving variant cannot damage the CUDA context of the
suite, and is marked xfail(strict=False): it reports (XPASS = fine on the GPU, XFAIL = needs attention) without ever stopping
`pytest -x`.  The file sorts last on purpose.  Promote a check to a normal test once it has passed on a GPU."""
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


@pytest.mark.xfail(strict=False, reason="not yet run on a GPU when written; XPASS means it is fine")
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variant_gives_identical_output_on_the_gpu(variant):
    assert run_worker("variant", *VARIANTS[variant]) == 0


@pytest.mark.xfail(strict=False, reason="not yet run on a GPU when written; XPASS means it is fine")
def test_late_entry_points_on_the_gpu():
    assert run_worker("features") == 0
