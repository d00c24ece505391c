"""bench.py cannot run without a GPU, but the pieces that turn library statistics into the JSON contract can: checked here against
statistics produced by the emulated library, so that a typo in bench.py is not first discovered on the GPU box."""
from __future__ import annotations

import json
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

import bench  # noqa: E402
from bcalm_b200 import api  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ to build the CPU emulation of the library")


@pytest.fixture(scope="module")
def emu():
    import build_emu
    path = build_emu.build()
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = path, None
    yield build_emu
    api.LIB_PATH, api._lib = saved


def _stats(env=None):
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        case = [c for c in json.load(f) if c["name"] == "reads_0_k15_a2"][0]
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        with api.Cdbg(k=case["k"], abundance_min=case["abundance_min"], device=0) as c:
            return c.build(case["reads"]).stats
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def test_dominant_kernel_groups_for_the_default_path(emu):
    st = _stats()
    name, ms, nbytes, groups, k2 = bench.dominant_kernel(st, st["n_bytes"])
    assert name in groups and ms == max(groups.values()) or abs(ms - max(groups.values())) < 1e-3
    assert "K1 extract (single pass)" in groups and "K2 count (fine partition + shared-memory tables)" in groups
    assert set(k2) == {"sk2p_hist", "sk2p_scatter", "sk2c_count"}       # stage 1 runs on super-k-mers by default
    I, S, R = st["n_instances"], st["n_solid"], st["n_superkmers"]
    assert k2["sk2c_count"][1] == I * 8 + S * 12 and k2["sk2p_scatter"][1] == 2 * I * 8 and k2["sk2c_count"][2] == R * 16 + S * 12
    st = _stats({"CDBG_S1": "kmer"})
    k2 = bench.dominant_kernel(st, st["n_bytes"])[4]
    assert set(k2) == {"k2p_hist", "k2p_scatter", "k2c_count"} and k2["k2c_count"][1] == st["n_instances"] * 8 + st["n_solid"] * 12
    # the committed ncu traffic figure is found under the group's name
    assert bench.load_traffic("K2 count (fine partition + shared-memory tables)") > 0


def test_dominant_kernel_groups_for_the_round_start_path(emu):
    st = _stats({"CDBG_S1": "kmer", "CDBG_K1": "twopass", "CDBG_K2": "l2", "CDBG_K5": "l2"})
    name, ms, nbytes, groups, k2 = bench.dominant_kernel(st, st["n_bytes"])
    assert "K1 extract (hist pass)" in groups and "K2 count (insert+emit, all bins)" in groups and k2 is None
    assert bench.load_traffic("K2 count (insert+emit, all bins)") > 0


def test_file_to_file_sample_runs_the_cli(emu, monkeypatch, tmp_path):
    g, n_reads = bench.workload(0.001)
    with api.Cdbg(k=bench.K, abundance_min=bench.AMIN, device=0) as c:
        ptr, n = c.synth_reads(g, bench.READ_LEN, n_reads, bench.ERR_PPM, bench.SEED)
        reads = c.copy_to_host(ptr, n)
    # point the helper at the emulated CLI
    fake_root = tmp_path / "root"
    (fake_root / "bcalm_b200").mkdir(parents=True)
    os.symlink(emu.CLI, fake_root / "bcalm_b200" / "bcalm")
    monkeypatch.setattr(bench, "ROOT", str(fake_root))
    out = bench.file_to_file_sample(reads, n_reads * bench.READ_LEN)
    assert out["value"] > 0 and out["fasta_bytes"] == n_reads * (bench.READ_LEN + 4) and out["cli_split"].startswith("wall s")


def test_numpy_read_generator_equals_the_device_generator(emu):
    """The reference arm generates its reads with numpy (no CUDA library in that process): the same bytes as synth_reads_kernel."""
    with api.Cdbg(k=bench.K, abundance_min=bench.AMIN, device=0) as c:
        for (g, L, first, n, seed) in ((10_000, 100, 0, 300, 42), (7_777, 150, 1234, 257, 43), (400, 100, 5, 3, 7)):
            ptr, nb = c.synth_reads(g, L, n, bench.ERR_PPM, seed, first_read=first)
            dev = c.copy_to_host(ptr, nb)
            assert (dev == bench.synth_reads_numpy(g, L, first, n, bench.ERR_PPM, seed, chunk=100)).all()


def test_workload_is_the_configuration_baseline_names():
    g, n_reads = bench.workload(1.0)
    assert (g, n_reads, bench.READ_LEN, bench.K, bench.AMIN) == (5_000_000, 1_500_000, 100, 31, 2)


def _run_bench_main(monkeypatch, capsys, argv):
    import torch
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **kw: real_empty(*a, **kw))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **kw: None)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    for v in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(v, raising=False)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_main_prints_the_contract_line_against_the_emulated_library(emu, monkeypatch, capsys):
    """The whole of bench.main() at a tiny scale, torch's CUDA calls stubbed: every key of the JSON contract is there and consistent."""
    out = _run_bench_main(monkeypatch, capsys, ["--steps", "2", "--warmup", "1", "--scale", "0.002", "--cpu-scale", "0.002", "--no-clocks",
                                                "--c3-scale", "0.00004", "--c3-steps", "1"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline", "file_to_file", "device_ms_per_step"):
        assert key in out, key
    assert out["value"] > 0 and out["n_gpus"] == 1 and out["steps"] == 2 and out["warmup"] == 1 and out["vs_baseline"] is None
    assert "workload" in out["config"] and "model" not in out["config"]
    e = out["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == out["config"]["input_bytes"] and e["d2h_bytes_per_step"] > 0
    r = out["roofline"]
    assert r["bound"] == "hbm" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["kernel"] in r["stage_ms"]
    assert out["gpu_launches"] > 0 and out["gpu_launches"] % 2 == 0
    cb = out["cpu_baseline"]
    assert cb["value"] and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"]
    # the emulated CUDA path reproduces the committed checksum of the oracle's graph of the same reads
    assert out["parity"] == "ok", out["parity_detail"]
    c3 = out["extra"]["c3"]
    assert c3["parity"] == "ok" and c3["value"] > 0 and c3["e2e"]["value"] > 0 and c3["scaling"] == "strong" and c3["workload"].startswith("C3")


def test_reference_arm_prints_its_line(emu, monkeypatch, capsys):
    out = _run_bench_main(monkeypatch, capsys, ["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-scale", "0.002"])
    assert out["impl"] == "reference" and out["value"] > 0 and out["e2e"]["value"] == out["value"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["cpu_baseline"]["kind"] in ("reference", "port")
    assert out["metric"] == "input Gbases/sec to unitigs at k=31" and out["unit"] == "Gbases/s"
    # like for like: the arm's config object is the one our arm prints for the same workload
    g, n = bench.workload(0.002)
    assert out["config"] == bench.c2_config(g, n)


def test_smoke_entry_point_against_the_emulated_library(emu, capsys):
    """__graft_entry__.smoke() is the first thing the driver runs on the GPU box: its own logic must not be what fails there."""
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke OK" in capsys.readouterr().out
