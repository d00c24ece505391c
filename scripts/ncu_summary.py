#!/usr/bin/env python
"""One line per profiled launch from `ncu --set full` reports: duration, DRAM bytes, pipe utilisation, IPC, occupancy, stalls.
    python scripts/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_summary.txt   (also prints a traffic JSON on the last line)"""
import csv, io, json, subprocess, sys

WANT = {
    "dur_us": "gpu__time_duration.sum", "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
    "dram%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l2%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1%": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm%": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "ipc": "sm__inst_executed.avg.per_cycle_elapsed", "occ%": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread", "smem_dyn_KB": "launch__shared_mem_per_block_dynamic", "winst": "smsp__inst_executed.sum",
    "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smem_pipe%": "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "bankconf": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "stall_lsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_bar": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_notsel": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "stall_lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


traffic = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        print(f"# {rep}: empty"); continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0]
        parts = [f"{name:26s}", f"grid={r[col['Grid Size']].replace(' ', '')}", f"block={r[col['Block Size']].replace(' ', '')}"]
        vals = {}
        for short, metric in WANT.items():
            if metric not in col:
                continue
            v = num(r[col[metric]])
            if v is None:
                continue
            u = units[col[metric]]
            if short.endswith("_MB"):
                v = v * SCALE.get(u, 1.0) / 1e6
            elif short == "dur_us":
                v = v * SCALE.get(u, 1e-3)
            elif short.endswith("_KB"):
                v = v * SCALE.get(u, 1.0) / 1e3
            vals[short] = v
            parts.append(f"{short}={v:.4g}")
        print(" ".join(parts))
        t = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "dur_us": 0.0})
        t["launches"] += 1
        t["dram_bytes"] += (vals.get("dram_rd_MB", 0) + vals.get("dram_wr_MB", 0)) * 1e6
        t["dur_us"] += vals.get("dur_us", 0)
print("TRAFFIC " + json.dumps(traffic))
