"""The CPU emulation of the CUDA runtime (tests/emu) is only worth trusting where it is at least as strict as a GPU.  These cases
feed it deliberate host-side mistakes -- an empty grid, dynamic shared memory without the opt-in, a host read before the stream
synchronisation, an event read too early, fresh device memory assumed zero -- and require that each one is caught."""
from __future__ import annotations

import os
import shutil
import signal
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("strict") / "strictness")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-w", "-I", os.path.join(HERE, "emu"), "-o", out,
                           os.path.join(HERE, "emu", "selftest", "strictness.cpp")])
    return out


def run(exe, case):
    return subprocess.run([exe, case], capture_output=True, text=True, timeout=60)


def test_a_correct_program_passes(exe):
    assert run(exe, "ok").returncode == 0


@pytest.mark.parametrize("case,needle", [("empty_grid", "empty grid"), ("smem_no_optin", "without the opt-in"), ("smem_optin_too_small", "without the opt-in")])
def test_launch_mistakes_abort(exe, case, needle):
    p = run(exe, case)
    assert p.returncode == -signal.SIGABRT and needle in p.stderr, (p.returncode, p.stderr)


@pytest.mark.parametrize("case", ["read_before_sync", "event_before_sync", "uninitialised"])
def test_host_side_mistakes_are_visible(exe, case):
    assert run(exe, case).returncode == 3
