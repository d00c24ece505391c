#!/usr/bin/env python
"""bench.py -- input Gbases/s from reads to final unitigs (+ links) at k=31.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale F] [--no-c3] ...

Main line, workload C2 (BASELINE.json configs[1]): synthetic 5 Mbp genome, 100 bp reads, 30x coverage (1.5 M reads, 150 Mbases),
1 % substitution errors, both strands, k=31, abundance-min 2.  One "step" = one full pass reads -> solid k-mers -> compaction ->
unitigs -> links over that read set.  At N > 1 the main line is WEAK scaling (5 Mbp of genome and 1.5 M reads per GPU, ONE graph
built by all ranks together, sharded by minimizer).

  value : reads already resident in HBM when the timed region starts (cdbg_build with a device pointer), results left on
          the device; wall time around K steps with a device synchronize (+ barrier) on both sides, max over ranks.
  e2e   : the same pass through the C ABI with HOST buffers: pinned host reads -> H2D -> all stages -> D2H of the unitig
          arrays (sequences, offsets, KC, links) inside the timed region.
  roofline : the dominant kernel group of the step (by CUDA-event time measured inside the library on its own stream),
          algorithmic bytes per SURVEY.md 8d / DESIGN.md 4; `kernels` lists the group's kernels one by one.
  parity : untimed.  N = 1: the checksum (bcalm_b200/checksum.py) of the fetched graph against tests/golden/c2_checksum.json, which
          tests/golden/make_c2_checksum.py computed from the C oracle's graph of the same reads.  N > 1: the merged shards against
          a single-GPU build of ALL reads on rank 0.  A mismatch prints "parity": "FAIL" and exits non-zero.
  extra.c3 : workload C3 (configs[2]: 250 Mbp genome, 150 bp reads, 30x = 7.5 Gbases) -- STRONG scaling: the same read set at
          every N, rank r taking an equal share; value, e2e, stage times and the same parity check (--no-c3 skips it).
  extra.c4 / extra.c5 : at 8 GPUs (or with --c4 / --c5) also C4 at full size (configs[3]: 3 Gbp genome, 90 Gbases; properties parity) and a
          scaled C5 (configs[4]: k=127, 250 bp reads, 3 Gbases; merged shards vs a single-GPU build).
  file_to_file : the same C2 workload through the CLI, FASTA file -> unitigs.fa (what the reference arm measures), N=1 only.
  cpu_baseline : the real reference binary (oracle/_ref/bcalm, kind "reference") -- or the C oracle (kind "port") when the
          binary is absent -- on the same C2 reads, all host cores (N=1, rank 0 only).

--impl reference times the reference's own CPU implementation (oracle/_ref/bcalm) on the SAME C2 read set (generated with numpy
by the same hash recipe as the device generator: no CUDA library is loaded in that arm), file -> file, one run per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
AMIN = 2
GENOME = 5_000_000
READ_LEN = 100
COVERAGE = 30
ERR_PPM = 10_000
SEED = 42
C3 = {"genome": 250_000_000, "read_len": 150, "coverage": 30, "seed": 43}
METRIC = "input Gbases/sec to unitigs at k=31"


def workload(scale: float):
    g = max(READ_LEN * 4, int(GENOME * scale))
    n_reads = g * COVERAGE // READ_LEN
    return g, n_reads


def c2_config(g: int, n_reads: int, world: int = 1) -> dict:
    """The `config` object of the JSON line: identical in both arms at N = 1."""
    n_bytes = (n_reads // world) * (READ_LEN + 1)
    wl = f"C2: synthetic {g} bp genome, {n_reads} x {READ_LEN} bp reads, {COVERAGE}x, 1% substitutions, k={K}, abundance-min {AMIN}"
    if world > 1:
        wl += (f"; weak scaling: 5 Mbp of genome and 1.5 M reads per GPU, ONE graph sharded by minimizer over {world} GPUs "
               f"(NCCL all-to-all of super-k-mers, junction records, glue and link records)")
    return {"workload": wl, "input_bytes": int(n_bytes),
            "l2_note": "input (151 MB per GPU) and its intermediate streams are larger than the 126 MB L2; no L2 flush needed"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- synthetic reads without CUDA

def _mix64(x):
    x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xff51afd7ed558ccd)
    x = x ^ (x >> np.uint64(33)); x = x * np.uint64(0xc4ceb9fe1a85ec53)
    return x ^ (x >> np.uint64(33))


def synth_reads_numpy(genome_len: int, read_len: int, first_read: int, n_reads: int, err_ppm: int, seed: int, chunk: int = 100_000) -> np.ndarray:
    """The same bytes as the device generator (`synth_reads_kernel`, bcalm_b200/csrc/kernels.cuh; equality is tested in
    tests/test_bench_helpers.py): genome base i = hash(seed, i), read r = hash(seed, r) -> start, strand; 1 % substitutions."""
    NT = np.frombuffer(b"ACTG", np.uint8)
    out = np.empty((n_reads, read_len + 1), np.uint8)
    stride, sd = np.uint64(read_len + 1), np.uint64(seed)
    j = np.arange(read_len, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        for c0 in range(0, n_reads, chunk):
            c1 = min(n_reads, c0 + chunk)
            r = (np.arange(c0, c1, dtype=np.uint64) + np.uint64(first_read))[:, None]
            hr = _mix64(sd * np.uint64(0xD1342543DE82EF95) + r + np.uint64(1))
            start = hr % np.uint64(genome_len - read_len + 1)
            rev = (hr >> np.uint64(63)) != 0
            gpos = np.where(rev, start + (np.uint64(read_len - 1) - j), start + j)
            b = (_mix64(sd ^ (gpos * np.uint64(0x9E3779B97F4A7C15))) >> np.uint64(61)).astype(np.uint32) & 3
            b = np.where(rev, b ^ 2, b)
            he = _mix64((sd + np.uint64(0x5851F42D4C957F2D)) ^ ((r * stride + j) * np.uint64(0xA24BAED4963EE407)))
            err = (he % np.uint64(1_000_000)).astype(np.uint32) < err_ppm
            b = np.where(err, (b + 1 + ((he >> np.uint64(40)) % np.uint64(3)).astype(np.uint32)) & 3, b)
            out[c0:c1, :read_len] = NT[b]
            out[c0:c1, read_len] = 10
    return out.reshape(-1)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read through NVML (the library behind nvidia-smi: the same
    clocks.sm / clocks_event_reasons counters) from a thread of this process every 20 ms.  Spawning `nvidia-smi -lms` instead
    stalls CUDA API calls of the measured process for 100-600 ms per poll on this driver (scripts/dev_sampler.py), which for
    a 6 ms step would measure the sampler, not the kernels."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.h = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the CUDA ordinal to the NVML handle through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(device).pci_bus_id if hasattr(torch.cuda.get_device_properties(device), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hh = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hh).bus == bus:
                        self.h = hh
                        break
        except Exception:
            self.h = None

    def _poll(self):
        while not self.stop_flag:
            try:
                self.samples.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                                     self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join()
        sm = [a for a, _ in self.samples]
        mask = 0
        for _, r in self.samples:
            mask |= int(r)
        reasons = sorted(name for bit, name in self.REASONS.items() if mask & bit)
        try:
            mx = float(self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                "source": "NVML (nvidia-smi's counters), 20 ms period, sampled during the timed region"}


def write_fasta_from_bytes(path: str, reads: np.ndarray, read_len: int):
    """reads: uint8 array of n_reads * (read_len + 1) bytes ('\\n' after each read) -> FASTA with '>r' headers."""
    n = reads.size // (read_len + 1)
    a = reads[: n * (read_len + 1)].reshape(n, read_len + 1)
    out = np.empty((n, read_len + 4), np.uint8)
    out[:, 0] = ord(">"); out[:, 1] = ord("r"); out[:, 2] = ord("\n")
    out[:, 3:] = a
    out.tofile(path)


def run_reference_binary(fasta: str, n_bases: int, cores: int):
    from oracle import oracle as orc
    text, dt, log = orc.run_ref(fasta, K, AMIN, nb_cores=cores, extra=["-verbose", "0"])
    return dt, text.count(">")


def cpu_baseline_sample(reads: np.ndarray, g: int, n_reads: int, cores: int):
    """The reference's CPU path on the same reads (all host cores); a scalar port only when the binary is absent."""
    from oracle import oracle as orc
    n_bases = n_reads * READ_LEN
    sample = f"{g} bp genome, {n_reads} x {READ_LEN} bp reads ({n_bases / 1e6:.1f} Mbases, {COVERAGE}x), k={K}, abundance-min {AMIN}"
    if orc.have_ref():
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "sample.fa")
            write_fasta_from_bytes(fa, reads, READ_LEN)
            dt, n_unitigs = run_reference_binary(fa, n_bases, cores)
        return {"value": n_bases / dt / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "reference",
                "sample": sample + f"; BCALM 2 binary -nb-cores {cores}, {dt:.2f} s wall, {n_unitigs} unitigs, file->file"}
    t0 = time.perf_counter()
    o = orc.run(reads, K, AMIN)
    dt = time.perf_counter() - t0
    return {"value": n_bases / dt / 1e9, "unit": "Gbases/s", "cores": 1, "kind": "port",
            "sample": sample + f"; C oracle single thread, {dt:.2f} s, {o.n_unitigs} unitigs"}


def file_to_file_sample(reads: np.ndarray, n_bases: int):
    """The same workload through the bcalm-compatible CLI, FASTA file -> <prefix>.unitigs.fa (process start, CUDA context, file read,
    parse, H2D, stages, D2H, formatting and write all inside): the variant SURVEY 8d asks for next to resident -> resident, and
    the like-for-like of the reference arm, which is file -> file too."""
    cli = os.path.join(ROOT, "bcalm_b200", "bcalm")
    if not os.path.exists(cli):
        return None
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "reads.fa")
        write_fasta_from_bytes(fa, reads, READ_LEN)
        best, log = None, []
        for _ in range(2):
            t0 = time.perf_counter()
            p = subprocess.run([cli, "-in", fa, "-kmer-size", str(K), "-abundance-min", str(AMIN), "-out", os.path.join(d, "out")],
                               capture_output=True, text=True, cwd=d)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                return {"value": None, "unit": "Gbases/s", "note": "CLI failed: " + p.stdout[-200:]}
            if best is None or dt < best:
                best, log = dt, [ln for ln in p.stdout.splitlines() if ln.startswith("wall s")]
        return {"value": n_bases / best / 1e9, "unit": "Gbases/s", "seconds": best, "fasta_bytes": os.path.getsize(fa),
                "cli_split": log[0] if log else None,
                "note": "bcalm_b200/bcalm -in reads.fa -> out.unitigs.fa, best of 2 runs, page cache warm; includes process start and CUDA context creation"}


def load_traffic(name: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel group from the committed `ncu --set full` captures."""
    for fn in ("r02_traffic.json", "r01b_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                return json.load(f)["groups"][name]["traffic_bytes_per_step"]
        except Exception:
            continue
    return None


def dominant_kernel(st: dict, n_bytes: int):
    """Pick the kernel group with the largest device time in the step and compute its algorithmic bytes (SURVEY 8d:
    the unit of K1/K2 is the k-mer instance at w = 8 bytes, whatever the implementation really moves).  Also returns the kernels of
    the K2 group with their own algorithmic bytes -- and, on the super-k-mer path, the bytes that format really has to move."""
    I, S, w = st["n_instances"], st["n_solid"], 8
    T = st["n_junction_records"]
    R = st.get("n_superkmers", 0)
    single = st["ms_k1_hist"] == 0
    k2_smem = st.get("ms_k2_table", 0) > 0
    groups = {}
    if single:
        groups["K1 extract (single pass)"] = (st["ms_k1_scatter"], n_bytes + I * w)
    else:
        groups["K1 extract (hist pass)"] = (st["ms_k1_hist"], n_bytes)
        groups["K1 extract (scatter pass)"] = (st["ms_k1_scatter"], n_bytes + I * w)
    groups["K2 count (fine partition + shared-memory tables)" if k2_smem else "K2 count (insert+emit, all bins)"] = (st["ms_k2"], I * w + S * (w + 4))
    groups["K3 end records (hist+scatter)"] = (st["ms_k3"], 2 * S * (w + 4) + T * 16)
    groups["K5 junction join"] = (st["ms_k5"], T * 16 + 2 * S * 4)
    groups["K6 list ranking"] = (st["ms_k6_rank"], 2 * S * 4 + 2 * S * 8 * 2)
    groups["K6/K8 numbering + spelling"] = (st["ms_k6_spell"], S * (w + 4) + 2 * S * 8 + st["total_nt"] + st["n_unitigs"] * 16)
    name = max(groups, key=lambda k_: groups[k_][0])
    ms, nbytes = groups[name]
    kernels = None
    if k2_smem and R:      # super-k-mer records (16 bytes) instead of k-mer instances
        kernels = {"sk2p_hist": (st["ms_k2_hist"], I * w, R * 16), "sk2p_scatter": (st["ms_k2_scatter"], 2 * I * w, 2 * R * 16),
                   "sk2c_count": (st["ms_k2_table"], I * w + S * (w + 4), R * 16 + S * (w + 4))}
    elif k2_smem:
        kernels = {"k2p_hist": (st["ms_k2_hist"], I * w, I * w), "k2p_scatter": (st["ms_k2_scatter"], 2 * I * w, 2 * I * w),
                   "k2c_count": (st["ms_k2_table"], I * w + S * (w + 4), I * w + S * (w + 4))}
    return name, ms, nbytes, {k_: round(v[0], 4) for k_, v in groups.items()}, kernels


def roofline_block(st: dict, n_bytes: int, with_traffic: bool):
    peak, peak_src = load_peaks()
    name, ms_k, bytes_k, groups, kernels = dominant_kernel(st, n_bytes)
    achieved = bytes_k / (ms_k / 1e3) / 1e9 if ms_k > 0 else 0.0
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": load_traffic(name) if with_traffic else None, "algorithmic_bytes": int(bytes_k), "kernel_ms": ms_k, "peak_source": peak_src,
            "stage_ms": groups,
            "shard_ms": None if not st.get("ms_x1") and not st.get("ms_glue") else {
                "X1 super-k-mer exchange (pack + all-to-all)": round(st["ms_x1"], 4), "X2 end-record exchange": round(st["ms_x2"], 4),
                "X3 link-message routing": round(st["ms_x3"], 4), "glue (inside K6/K8 numbering + spelling)": round(st["ms_glue"], 4),
                "link-record routing": round(st["ms_xlink"], 4), "count stage": round(st["ms_count"], 4), "compact stage": round(st["ms_compact"], 4),
                "link stage": round(st["ms_link"], 4), "pieces_rank0": int(st.get("n_pieces", 0))},
            # per kernel: `achieved` / `frac` on the bytes the kernel's data format really moves (super-k-mer records are ~4x smaller than
            # the 8-byte k-mer instances of SURVEY 8d's accounting, which `algorithmic_bytes` keeps for comparison with round 1)
            "kernels": None if not kernels else {k_: {"ms": round(m_, 4), "algorithmic_bytes": int(b_), "format_bytes": int(f_),
                                                      "achieved": round(f_ / (m_ / 1e3) / 1e9, 1) if m_ > 0 else None,
                                                      "frac": round(f_ / (m_ / 1e3) / 1e9 / peak, 4) if m_ > 0 else None} for k_, (m_, b_, f_) in kernels.items()}}


# ---------------------------------------------------------------------------------------------- parity (untimed)

CHECK_KEYS = ("n_unitigs", "total_nt", "sum_kc", "n_links", "nodes", "links", "bad_overlaps")


def sharded_parity(c, world, rank, local_rank, all_reads_fn, k, amin, dist):
    """Merged shards of the LAST build of `c` vs a single-GPU build of all reads on rank 0.  Returns a dict on rank 0, None elsewhere."""
    from bcalm_b200.api import Cdbg, merge_shards
    from bcalm_b200.checksum import graph_checksum
    g = c.unitigs()
    gathered = [None] * world
    dist.gather_object(g, gathered if rank == 0 else None, dst=0)
    out = None
    if rank == 0:
        merged = graph_checksum(merge_shards(gathered), device=_dev(local_rank))
        with Cdbg(k=k, abundance_min=amin, device=local_rank) as single:
            from bcalm_b200.api import MEM_DEVICE
            ptr, n = all_reads_fn(single)
            single.build_raw(ptr, n, MEM_DEVICE, fetch=True)
            ref = graph_checksum(single.unitigs(copy=False), device=_dev(local_rank))
        ok = merged == ref and merged["bad_overlaps"] == 0
        out = {"ok": bool(ok), "check": f"checksum of the merged shards of {world} GPUs == checksum of a single-GPU build of all reads (rank 0)",
               "checksum": merged, **({} if ok else {"single_gpu": ref})}
        try:
            import torch
            torch.cuda.empty_cache()        # the checker's tensors must not stay in the way of the next workload
        except Exception:
            pass
    dist.barrier()
    return out


def golden_parity(g, path_key: str):
    """N = 1: checksum of the fetched graph vs the committed checksum of the C oracle's graph of the same reads."""
    from bcalm_b200.checksum import graph_checksum
    cs = graph_checksum(g, device=_dev(0))
    try:
        with open(os.path.join(ROOT, "tests", "golden", "c2_checksum.json")) as f:
            gold = json.load(f)[path_key]
    except Exception:
        gold = None
    if gold is None:
        ok = cs["bad_overlaps"] == 0
        return {"ok": bool(ok), "check": "no committed checksum for this scale: link overlaps only", "checksum": cs}
    ok = all(cs[k_] == gold[k_] for k_ in CHECK_KEYS)
    return {"ok": bool(ok), "check": "checksum == tests/golden/c2_checksum.json (the C oracle's graph of the same reads; tests/golden/make_c2_checksum.py)",
            "checksum": cs, **({} if ok else {"expected": gold})}


def _dev(local_rank: int) -> str:
    try:
        import torch
        return f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu"
    except Exception:
        return "cpu"


class Watchdog:
    """A collective that one rank never enters (a rank failed before it) blocks the others inside a C call for ever.  The extra
    blocks run under this timer: when it fires, rank 0 prints the line it has so far and every rank leaves."""

    def __init__(self, seconds: float, rank: int, payload):
        self.t = threading.Timer(seconds, self._fire)
        self.t.daemon = True
        self.rank, self.payload, self.seconds = rank, payload, seconds

    def _fire(self):
        if self.rank == 0:
            out = self.payload()
            if out is not None:
                print(json.dumps(out), flush=True)
        os._exit(0 if self.rank == 0 else 0)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.t.cancel()


# ---------------------------------------------------------------------------------------------- C3 block (strong scaling)

def run_c3(args, rank, local_rank, world, barrier, dist, torch):
    from bcalm_b200.api import Cdbg, MEM_DEVICE, MEM_HOST
    G, L = int(C3["genome"] * args.c3_scale), C3["read_len"]
    n_total = G * C3["coverage"] // L
    n_reads = n_total // world
    first = rank * n_reads
    if rank == world - 1:
        n_reads = n_total - first
    c = Cdbg(k=K, abundance_min=AMIN, device=local_rank, rank=rank, world_size=world)
    if world > 1:
        c.shard_init_torch()
    ptr, n_bytes = c.synth_reads(G, L, n_reads, ERR_PPM, C3["seed"], first_read=first)
    steps, warm = args.c3_steps, 1
    for _ in range(warm):
        c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
    barrier()
    t0 = time.perf_counter()
    agg = None
    for _ in range(steps):
        v, s = c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
        d = s.as_dict()
        agg = d if agg is None else {k_: (agg[k_] + d[k_] if k_.startswith("ms_") else d[k_]) for k_ in d}
    barrier()
    dt = (time.perf_counter() - t0) / steps
    st = {k_: (agg[k_] / steps if k_.startswith("ms_") else agg[k_]) for k_ in agg}
    # e2e: pinned host reads
    host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    hnp[:] = c.copy_to_host(ptr, n_bytes)
    c.build_raw(hnp.ctypes.data, n_bytes, MEM_HOST, fetch=True)
    barrier()
    t1 = time.perf_counter()
    for _ in range(steps):
        v, s2 = c.build_raw(hnp.ctypes.data, n_bytes, MEM_HOST, fetch=True)
    barrier()
    dt_e2e = (time.perf_counter() - t1) / steps
    U, nt, nl = v.n_unitigs, v.total_nt, v.n_links
    d2h = nt + (U + 1) * 8 + U * 8 + nl * 12 + (U + 1) * 8
    del host, hnp
    if world > 1:
        tt = torch.tensor([dt, dt_e2e], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = tt.tolist()
        par = sharded_parity(c, world, rank, local_rank, lambda single: single.synth_reads(G, L, n_total, ERR_PPM, C3["seed"]), K, AMIN, dist)
    else:
        from bcalm_b200.checksum import graph_checksum
        cs = graph_checksum(c.unitigs(copy=False), device=_dev(local_rank))
        ok = cs["bad_overlaps"] == 0 and cs["total_nt"] - cs["n_unitigs"] * (K - 1) == int(st["n_solid"])
        par = {"ok": bool(ok), "check": "every link overlaps by k-1 and the unitigs spell each solid k-mer once (sum(LN-k+1) == solid k-mers); "
                                        "the checksum is what the N > 1 runs must reproduce", "checksum": cs}
    c.close()
    if rank != 0:
        return None
    n_bases = n_total * L
    return {"workload": f"C3: synthetic {G} bp genome, {n_total} x {L} bp reads, {C3['coverage']}x, 1% substitutions, k={K}, abundance-min {AMIN}; "
                        + ("one GPU" if world == 1 else f"strong scaling: the same read set split over {world} GPUs, ONE graph sharded by minimizer"),
            "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warm, "value": n_bases / dt / 1e9, "unit": "Gbases/s", "ms_per_step": dt * 1e3,
            "e2e": {"value": n_bases / dt_e2e / 1e9, "unit": "Gbases/s", "ms_per_step": dt_e2e * 1e3, "h2d_bytes_per_step": int(n_bytes), "d2h_bytes_per_step": int(d2h),
                    "zero_copy": bool(s2.zero_copy)},
            "roofline": roofline_block(st, n_bytes, False), "device_ms_per_step_rank0": round(st["ms_count"] + st["ms_compact"] + st["ms_link"], 3),
            "solid_kmers_rank0": int(st["n_solid"]), "superkmers_rank0": int(st.get("n_superkmers", 0)), "k2_splits_rank0": int(st.get("n_k2_splits", 0)),
            "parity": "ok" if par["ok"] else "FAIL", "parity_detail": par}


def run_c4(args, rank, local_rank, world, barrier, dist, torch):
    """BASELINE.json configs[3]: 3 Gbp genome, 6e8 x 150 bp reads (90 Gbases), k=31, sharded over the GPUs of the box.  The input does
    not fit one GPU, so parity is by size-independent properties summed over the ranks."""
    from bcalm_b200.api import Cdbg, MEM_DEVICE
    G, L = int(3_000_000_000 * args.c4_scale), 150
    n_total = G * 30 // L
    n_reads = n_total // world
    first = rank * n_reads
    if rank == world - 1:
        n_reads = n_total - first
    c = Cdbg(k=K, abundance_min=AMIN, device=local_rank, rank=rank, world_size=world)
    if world > 1:
        c.shard_init_torch()
    ptr, n_bytes = c.synth_reads(G, L, n_reads, ERR_PPM, 44, first_read=first)
    c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)          # warm-up (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.c4_steps):
        v, s = c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
    barrier()
    dt = (time.perf_counter() - t0) / args.c4_steps
    st = s.as_dict()
    tt = torch.tensor([float(st["n_solid"]), float(st["total_nt"] - st["n_unitigs"] * (K - 1)), float(st["n_unitigs"]), float(st["n_links"]), float(st["n_pieces"])],
                      device=_dev(local_rank), dtype=torch.float64)
    tmax = torch.tensor([dt], device=_dev(local_rank), dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt = float(tmax.item())
    solid, spelled, U, links, pieces = [int(x) for x in tt.tolist()]
    c.close()
    if rank != 0:
        return None
    n_bases = n_total * L
    ok = solid == spelled and solid > 0
    return {"workload": f"C4: synthetic {G} bp genome, {n_total} x {L} bp reads ({n_bases / 1e9:.1f} Gbases), 30x, 1% substitutions, k={K}, abundance-min {AMIN}; "
                        f"sharded over {world} GPUs, reads resident in HBM",
            "n_gpus": world, "steps": args.c4_steps, "value": n_bases / dt / 1e9, "unit": "Gbases/s", "ms_per_step": dt * 1e3,
            "solid_kmers": solid, "unitigs": U, "links": links, "pieces_before_glue": pieces, "stage1_waves": int(st.get("sk_waves", 0)),
            "roofline": roofline_block(st, n_bytes, False),
            "parity": "ok" if ok else "FAIL",
            "parity_detail": {"ok": bool(ok), "check": "properties only (no single GPU holds this input): the unitigs of all ranks spell every solid k-mer exactly once "
                                                      "(sum(LN-k+1) == solid k-mers, summed over the ranks)", "solid": solid, "spelled": spelled}}


def run_c5(args, rank, local_rank, world, barrier, dist, torch):
    """BASELINE.json configs[4] scaled to what the multi-word path holds in HBM: k=127 (4-word k-mers), 250 bp reads, 30x, sharded by
    k-mer hash.  Parity: the merged shards against a single-GPU build of all reads on rank 0 -- (unitigs, nt, KC sum, links) and the
    multiset of (LN, KC) pairs (graph_checksum handles one-word k-mers only)."""
    from bcalm_b200.api import Cdbg, MEM_DEVICE, merge_shards
    k5 = 127
    G, L = int(args.c5_genome), 250
    n_total = G * 30 // L
    n_reads = n_total // world
    first = rank * n_reads
    if rank == world - 1:
        n_reads = n_total - first
    c = Cdbg(k=k5, abundance_min=AMIN, device=local_rank, rank=rank, world_size=world)
    if world > 1:
        c.shard_init_torch()
    ptr, n_bytes = c.synth_reads(G, L, n_reads, ERR_PPM, 45, first_read=first)
    c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.c5_steps):
        v, s = c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=True)
    barrier()
    dt = (time.perf_counter() - t0) / args.c5_steps

    def lite(g):
        ln = np.diff(np.asarray(g.seq_off, dtype=np.int64))
        kc = np.asarray(g.kc, dtype=np.uint64)
        order = np.lexsort((kc, ln))
        h = np.uint64(1469598103934665603)
        with np.errstate(over="ignore"):
            mixed = (ln[order].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (kc[order] * np.uint64(0xC2B2AE3D27D4EB4F))
            acc = np.bitwise_xor.reduce(mixed * (np.arange(1, len(mixed) + 1, dtype=np.uint64) | np.uint64(1))) if len(mixed) else np.uint64(0)
        return {"n_unitigs": int(len(ln)), "total_nt": int(ln.sum()), "sum_kc": int(kc.sum()), "n_links": int(len(g.links)), "ln_kc_multiset": int(acc ^ h)}

    par = None
    g = c.unitigs()
    if world > 1:
        tmax = torch.tensor([dt], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())
        gathered = [None] * world
        dist.gather_object(g, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            merged = lite(merge_shards(gathered))
            with Cdbg(k=k5, abundance_min=AMIN, device=local_rank) as single:
                p1, n1 = single.synth_reads(G, L, n_total, ERR_PPM, 45)
                single.build_raw(p1, n1, MEM_DEVICE, fetch=True)
                ref = lite(single.unitigs(copy=False))
            par = {"ok": merged == ref, "check": f"(unitigs, nt, KC sum, links, multiset of (LN, KC)) of the merged shards of {world} GPUs == single-GPU build of all reads",
                   "summary": merged, **({} if merged == ref else {"single_gpu": ref})}
        dist.barrier()
    else:
        cs = lite(g)
        st1 = s.as_dict()
        ok = cs["total_nt"] - cs["n_unitigs"] * (k5 - 1) == int(st1["n_solid"])
        par = {"ok": bool(ok), "check": "the unitigs spell every solid k-mer once; the summary is what the N > 1 runs must reproduce", "summary": cs}
    st = s.as_dict()
    c.close()
    if rank != 0:
        return None
    n_bases = n_total * L
    return {"workload": f"C5 (scaled): synthetic {G} bp genome, {n_total} x {L} bp reads ({n_bases / 1e9:.2f} Gbases), 30x, 1% substitutions, k={k5} (4-word k-mers), "
                        f"abundance-min {AMIN}; " + ("one GPU" if world == 1 else f"sharded by k-mer hash over {world} GPUs"),
            "n_gpus": world, "steps": args.c5_steps, "value": n_bases / dt / 1e9, "unit": "Gbases/s", "ms_per_step": dt * 1e3,
            "note": "timed through the C ABI with the reads resident in HBM and the unitigs fetched to the host",
            "solid_kmers_rank0": int(st["n_solid"]), "parity": "ok" if par["ok"] else "FAIL", "parity_detail": par}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="workload scale (1.0 = C2)")
    ap.add_argument("--cpu-scale", type=float, default=1.0, help="CPU baseline / reference arm workload, fraction of C2 (1.0 = the same reads as our arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-to-file", action="store_true", help="skip the CLI file -> file sample")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region (debugging)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed parity checks")
    ap.add_argument("--no-c3", action="store_true", help="skip the extra C3 block")
    ap.add_argument("--c3-scale", type=float, default=1.0)
    ap.add_argument("--c3-steps", type=int, default=3)
    ap.add_argument("--c4", action="store_true", help="also run C4 (3 Gbp genome, 90 Gbases; needs the HBM of 8 GPUs): on by default at --gpus 8")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--c4-scale", type=float, default=1.0)
    ap.add_argument("--c4-steps", type=int, default=1)
    ap.add_argument("--c5", action="store_true", help="also run a scaled C5 (k=127, 250 bp reads): on by default at --gpus 8")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--c5-genome", type=float, default=50_000_000)
    ap.add_argument("--c5-steps", type=int, default=2)
    ap.add_argument("--c3-timeout", type=float, default=420.0, help="seconds after which a hung C3 block is abandoned (the main line is still printed)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm: the reference's own CPU path
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as orc
        g, n_reads = workload(args.cpu_scale)
        n_bases = n_reads * READ_LEN
        reads = synth_reads_numpy(g, READ_LEN, 0, n_reads, ERR_PPM, SEED)      # no CUDA library in this arm
        times = []
        kind = "reference" if orc.have_ref() else "port"
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "sample.fa")
            write_fasta_from_bytes(fa, reads, READ_LEN)
            for it in range(args.warmup + args.steps):
                if kind == "reference":
                    dt, _ = run_reference_binary(fa, n_bases, cores)
                else:
                    t0 = time.perf_counter(); orc.run(reads, K, AMIN); dt = time.perf_counter() - t0
                if it >= args.warmup:
                    times.append(dt)
        ms = float(np.mean(times)) * 1e3
        val = n_bases / (ms / 1e3) / 1e9
        sample = f"{g} bp genome, {n_reads} x {READ_LEN} bp reads ({n_bases / 1e6:.1f} Mbases, {COVERAGE}x) per step, file -> file, BCALM 2 -nb-cores {cores}"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "Gbases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": c2_config(g, n_reads),
            "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from bcalm_b200.api import Cdbg, MEM_DEVICE, MEM_HOST

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local_rank)

    # weak scaling: the genome grows with the number of GPUs (5 Mbp per GPU); ONE graph is built by all ranks together, every
    # rank contributing an equal share of the reads (sharded by minimizer, NCCL all-to-alls inside the library)
    g, n_reads_total = workload(args.scale * world)
    n_reads = n_reads_total // world
    n_bases = n_reads * READ_LEN
    c = Cdbg(k=K, abundance_min=AMIN, device=local_rank, rank=rank, world_size=world)
    if world > 1:
        c.shard_init_torch()
    ptr, n_bytes = c.synth_reads(g, READ_LEN, n_reads, ERR_PPM, SEED, first_read=rank * n_reads)

    # ---- value: reads resident in HBM
    for _ in range(args.warmup):
        c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0 and not args.no_clocks:
        sampler.start()
    t0 = time.perf_counter()
    agg = None
    for _ in range(args.steps):
        v, s = c.build_raw(ptr, n_bytes, MEM_DEVICE, fetch=False)
        d = s.as_dict()
        agg = d if agg is None else {k_: (agg[k_] + d[k_] if k_.startswith("ms_") else d[k_]) for k_ in d}
    barrier()
    dt = time.perf_counter() - t0
    st = {k_: (agg[k_] / args.steps if k_.startswith("ms_") else agg[k_]) for k_ in agg}

    # ---- e2e: pinned host reads -> H2D -> stages -> D2H of the unitig arrays
    host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    hnp[:] = c.copy_to_host(ptr, n_bytes)
    for _ in range(max(1, args.warmup)):
        c.build_raw(hnp.ctypes.data, n_bytes, MEM_HOST, fetch=True)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        v, s2 = c.build_raw(hnp.ctypes.data, n_bytes, MEM_HOST, fetch=True)
    barrier()
    dt_e2e = time.perf_counter() - t1
    clocks = sampler.stop() if rank == 0 else None
    U, nt, nl = v.n_unitigs, v.total_nt, v.n_links
    d2h = nt + (U + 1) * 8 + U * 8 + nl * 12 + (U + 1) * 8

    if world > 1:
        tt = torch.tensor([dt, dt_e2e], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = tt.tolist()

    # ---- parity of what was just timed (untimed)
    par = None
    if not args.no_parity:
        if world > 1:
            par = sharded_parity(c, world, rank, local_rank, lambda single: single.synth_reads(g, READ_LEN, n_reads * world, ERR_PPM, SEED), K, AMIN, dist)
        else:
            par = golden_parity(c.unitigs(copy=False), f"scale_{args.scale:g}")

    out = None
    if rank == 0:
        total_bases = n_bases * world
        out = {
            "metric": METRIC,
            "value": total_bases / dt * args.steps / 1e9, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": c2_config(g, n_reads * world, world),
            "e2e": {"value": total_bases / dt_e2e * args.steps / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": int(n_bytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": dt_e2e / args.steps * 1e3,
                    "zero_copy": bool(s2.zero_copy),
                    "note": "zero_copy: the pinned host reads are not copied first; the extraction kernel's TMA loads read them in place over PCIe, so the "
                            "h2d bytes cross the bus inside K1" if s2.zero_copy else "host reads copied H2D (cudaMemcpyAsync), then the stages"},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "device_ms_per_step": round(st["ms_count"] + st["ms_compact"] + st["ms_link"], 4),   # CUDA events on the library's stream (rank 0)
            "roofline": roofline_block(st, n_bytes, world == 1),
            "clocks": clocks,
            "parity": None if par is None else ("ok" if par["ok"] else "FAIL"),
            "parity_detail": par,
            "extra": {"unitigs_rank0": int(st["n_unitigs"]), "links_rank0": int(st["n_links"]), "solid_kmers_rank0": int(st["n_solid"]),
                      "superkmers_rank0": int(st.get("n_superkmers", 0)), "minimizer_size": int(st.get("sk_m", 0))},
        }
        if not args.no_file_to_file and world == 1:
            try:
                out["file_to_file"] = file_to_file_sample(c.copy_to_host(ptr, n_bytes), n_bases)
            except Exception as e:  # a reported extra, never a reason to lose the line
                out["file_to_file"] = {"value": None, "unit": "Gbases/s", "note": f"failed: {e}"}
        if not args.no_cpu_baseline and world == 1:   # the CPU leg runs on rank 0 at N=1 only
            try:
                gc, nc = workload(args.cpu_scale)
                reads_c = c.copy_to_host(ptr, n_bytes) if (gc, nc) == (g, n_reads) else synth_reads_numpy(gc, READ_LEN, 0, nc, ERR_PPM, SEED)
                out["cpu_baseline"] = cpu_baseline_sample(reads_c, gc, nc, cores)
            except Exception as e:  # never lose the GPU line because the CPU leg failed
                out["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}
    c.close()
    del host, hnp

    # ---- extra: C3, strong scaling
    if not args.no_c3:
        def partial():
            if out is not None:
                out["extra"]["c3"] = {"value": None, "parity": None, "note": f"abandoned after {args.c3_timeout:.0f} s (a rank failed or hung)"}
            return out
        with Watchdog(args.c3_timeout, rank, partial):
            try:
                c3 = run_c3(args, rank, local_rank, world, barrier, dist, torch)
            except Exception as e:
                c3 = {"value": None, "parity": None, "note": f"failed: {e}"}
                if world > 1:                     # the other ranks are inside a collective this rank will never enter
                    if rank == 0:
                        out["extra"]["c3"] = c3
                        print(json.dumps(out), flush=True)
                    time.sleep(5 if rank == 0 else 30)
                    os._exit(0)
        if rank == 0:
            out["extra"]["c3"] = c3

    # ---- extra (opt-in): C4 at full size and a scaled C5, each under its own watchdog
    want_c4 = (args.c4 or (world >= 8 and args.impl == "ours")) and not args.no_c4
    want_c5 = (args.c5 or (world >= 8 and args.impl == "ours")) and not args.no_c5
    if want_c5 and not args.c5:
        args.c5_genome = 100_000_000
    for name, flag, fn in (("c4", want_c4, run_c4), ("c5", want_c5, run_c5)):
        if not flag:
            continue
        def partial(name=name):
            if out is not None:
                out["extra"][name] = {"value": None, "parity": None, "note": f"abandoned after {args.c3_timeout:.0f} s (a rank failed or hung)"}
            return out
        with Watchdog(args.c3_timeout, rank, partial):
            try:
                res = fn(args, rank, local_rank, world, barrier, dist, torch)
            except Exception as e:
                res = {"value": None, "parity": None, "note": f"failed: {e}"}
                if world > 1:
                    if rank == 0:
                        out["extra"][name] = res
                        print(json.dumps(out), flush=True)
                    time.sleep(5 if rank == 0 else 30)
                    os._exit(0)
        if rank == 0:
            out["extra"][name] = res

    fail = False
    if rank == 0:
        print(json.dumps(out))
        fail = out["parity"] == "FAIL" or any(isinstance(out["extra"].get(n_), dict) and out["extra"][n_].get("parity") == "FAIL" for n_ in ("c3", "c4", "c5"))
    if world > 1:
        dist.destroy_process_group()
    if fail:
        sys.exit(1)


if __name__ == "__main__":
    main()
