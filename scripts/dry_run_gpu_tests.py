#!/usr/bin/env python
"""TEST INFRASTRUCTURE (manual): run the `-m gpu` test files here, without a GPU, against the CPU emulation of the library
(tests/emu) -- to find mistakes in the TESTS (a typo, a changed statistics field, a wrong fixture) before they cost a GPU call.
It proves nothing about the GPU: the real run is `python -m pytest tests -m gpu` on the B200 box.

    python scripts/dry_run_gpu_tests.py [pytest args ...]      # default: parity, CLI and the not-yet-validated variants
The sharded GPU tests (torchrun + NCCL) cannot run this way; tests/test_emu_sharded.py covers that path on emulated ranks.
tests/test_gpu_full_scale.py works too but takes ~4 minutes (the C2 read set at 1/100000 of the GPU's speed).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")]

import build_emu  # noqa: E402
import bcalm_b200.api as api  # noqa: E402
import bcalm_b200.build as build  # noqa: E402
import pytest  # noqa: E402

api.LIB_PATH = build_emu.build()
api._lib = None
build.build_cli = lambda force=False: build_emu.CLI            # the unchanged CLI source linked against the emulated library
os.environ["CDBG_GPU_TEST_LIB"] = api.LIB_PATH                  # for the subprocess worker of the variant tests
args = sys.argv[1:] or ["tests/test_gpu_parity.py", "tests/test_gpu_cli.py", "tests/test_gpu_variants.py"]
# two tests need the real thing: a torch CUDA tensor, and pinned memory that is told apart from pageable memory
SKIP = "not unaligned_falls_back_to_copy and not pinned_host_input_is_read_in_place"
sys.exit(pytest.main([*args, "-q", "-m", "gpu", "-k", SKIP, "-p", "no:cacheprovider"]))
