#!/usr/bin/env python
"""bench.py -- input Gbases/s from reads to final unitigs (+ links) at k=31.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale F]

Workload (BASELINE.json configs[1], "C2"): synthetic 5 Mbp genome, 100 bp reads, 30x coverage (1.5 M reads, 150 Mbases),
1 % substitution errors, both strands, k=31, abundance-min 2.  One "step" = one full pass reads -> solid k-mers ->
compaction -> unitigs -> links over that read set.

  value : reads already resident in HBM when the timed region starts (cdbg_build with a device pointer), results left on
          the device; wall time around K steps with a device synchronize on both sides, max over ranks.
  e2e   : the same pass through the C ABI with HOST buffers: pinned host reads -> H2D -> all stages -> D2H of the unitig
          arrays (sequences, offsets, KC, links) inside the timed region.
  roofline : the dominant kernel group of the step (by CUDA-event time measured inside the library on its own stream),
          algorithmic bytes per SURVEY.md 8d / DESIGN.md; for the K2 group also its three kernels one by one (`k2_kernels`).
  file_to_file : the same workload through the CLI, FASTA file -> unitigs.fa (what the reference arm measures), N=1 only.
  cpu_baseline : the real reference binary (oracle/_ref/bcalm, kind "reference") -- or the C oracle (kind "port") when the
          binary is absent -- on a bounded sample of the same workload, all host cores.

--impl reference times the reference's own CPU implementation (oracle/_ref/bcalm) on a bounded sample per step.
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


def workload(scale: float):
    g = max(READ_LEN * 4, int(GENOME * scale))
    n_reads = g * COVERAGE // READ_LEN
    return g, n_reads


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def cpu_baseline_sample(c, scale: float, cores: int):
    """Bounded sample of the same workload (same read length / coverage / error rate, smaller genome) on the host cores."""
    from oracle import oracle as orc
    from bcalm_b200.api import Cdbg
    g, n_reads = workload(scale)
    with Cdbg(k=K, abundance_min=AMIN) as cs:
        ptr, n = cs.synth_reads(g, READ_LEN, n_reads, ERR_PPM, SEED + 1)
        reads = cs.copy_to_host(ptr, n)
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
    host parse, H2D, stages, D2H, formatting and write all inside): the variant SURVEY 8d asks for next to resident -> resident, and
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
    for fn in ("r01b_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                return json.load(f)["groups"][name]["traffic_bytes_per_step"]
        except Exception:
            continue
    return None


def dominant_kernel(st: dict, n_bytes: int):
    """Pick the kernel group with the largest device time in the step and compute its algorithmic bytes.
    Also returns, for the K2 group, the three kernels it consists of with their own algorithmic bytes."""
    I, S, w = st["n_instances"], st["n_solid"], 8
    T = st["n_junction_records"]
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
    if k2_smem:
        kernels = {"k2p_hist": (st["ms_k2_hist"], I * w), "k2p_scatter": (st["ms_k2_scatter"], 2 * I * w),
                   "k2c_count": (st["ms_k2_table"], I * w + S * (w + 4))}
    return name, ms, nbytes, {k_: round(v[0], 4) for k_, v in groups.items()}, kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="workload scale (1.0 = C2)")
    ap.add_argument("--cpu-scale", type=float, default=0.5, help="bounded sample for the CPU baseline, fraction of C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-to-file", action="store_true", help="skip the CLI file -> file sample")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region (debugging)")
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
        from bcalm_b200.api import Cdbg
        g, n_reads = workload(args.cpu_scale)
        n_bases = n_reads * READ_LEN
        with Cdbg(k=K, abundance_min=AMIN, device=local_rank) as cs:   # only to generate the same synthetic reads
            ptr, n = cs.synth_reads(g, READ_LEN, n_reads, ERR_PPM, SEED + 1)
            reads = cs.copy_to_host(ptr, n)
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
        sample = f"{g} bp genome, {n_reads} x {READ_LEN} bp reads ({n_bases / 1e6:.1f} Mbases, {COVERAGE}x) per step"
        print(json.dumps({
            "impl": "reference", "metric": "input Gbases/sec to unitigs at k=31", "value": val, "unit": "Gbases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C2-like sample: {sample}, k={K}, abundance-min {AMIN}, BCALM 2 -nb-cores {cores}"},
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

    if rank == 0:
        peak, peak_src = load_peaks()
        name, ms_k, bytes_k, groups, k2_kernels = dominant_kernel(st, n_bytes)
        achieved = bytes_k / (ms_k / 1e3) / 1e9 if ms_k > 0 else 0.0
        total_bases = n_bases * world
        out = {
            "metric": "input Gbases/sec to unitigs at k=31",
            "value": total_bases / dt * args.steps / 1e9, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"C2: synthetic {g} bp genome, {n_reads * world} x {READ_LEN} bp reads, {COVERAGE}x, 1% substitutions, k={K}, "
                                   f"abundance-min {AMIN}" + ("" if world == 1 else f"; weak scaling: 5 Mbp of genome and 1.5 M reads per GPU, ONE graph "
                                   f"sharded by minimizer over {world} GPUs (NCCL all-to-all of k-mers, junction records, glue and link records)"),
                       "input_bytes": n_bytes, "l2_note": "input (151 MB) and k-mer stream (840 MB) are larger than the 126 MB L2; no L2 flush needed",
                       "unitigs_rank0": int(st["n_unitigs"]), "links_rank0": int(st["n_links"]), "solid_kmers_rank0": int(st["n_solid"])},
            "e2e": {"value": total_bases / dt_e2e * args.steps / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": int(n_bytes), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": dt_e2e / args.steps * 1e3,
                    "zero_copy": bool(s2.zero_copy),
                    "note": "zero_copy: the pinned host reads are not copied first; the extraction kernel's TMA loads read them in place over PCIe, so the "
                            "h2d bytes cross the bus inside K1" if s2.zero_copy else "host reads copied H2D (cudaMemcpyAsync), then the stages"},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "device_ms_per_step": round(st["ms_count"] + st["ms_compact"] + st["ms_link"], 4),   # CUDA events on the library's stream (rank 0)
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(name) if world == 1 else None, "algorithmic_bytes": int(bytes_k), "kernel_ms": ms_k, "peak_source": peak_src,
                         "stage_ms": groups,
                         "k2_kernels": None if not k2_kernels else {k_: {"ms": round(m_, 4), "algorithmic_bytes": int(b_), "achieved": round(b_ / (m_ / 1e3) / 1e9, 1) if m_ > 0 else None,
                                                                        "frac": round(b_ / (m_ / 1e3) / 1e9 / peak, 4) if m_ > 0 else None} for k_, (m_, b_) in k2_kernels.items()}},
            "clocks": clocks,
        }
        if not args.no_file_to_file and world == 1:
            try:
                out["file_to_file"] = file_to_file_sample(c.copy_to_host(ptr, n_bytes), n_bases)
            except Exception as e:  # a reported extra, never a reason to lose the line
                out["file_to_file"] = {"value": None, "unit": "Gbases/s", "note": f"failed: {e}"}
        if not args.no_cpu_baseline and world == 1:   # the CPU leg runs on rank 0 at N=1 only
            try:
                out["cpu_baseline"] = cpu_baseline_sample(c, args.cpu_scale, cores)
            except Exception as e:  # never lose the GPU line because the CPU leg failed
                out["cpu_baseline"] = {"value": None, "unit": "Gbases/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(out))
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
