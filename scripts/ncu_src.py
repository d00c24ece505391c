# dev helper: per-source-line instruction counts and stall samples from an ncu report (needs -lineinfo + --import-source on)
#   python scripts/ncu_src.py report.ncu-rep <kernel regex> [min percent]
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "-k", "regex:" + kern, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows, hdr, cur = [], None, ""
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0] in ("File Path", "File Name"): cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr and len(row) == len(hdr) and row[2] == "-":     # the per-source-line aggregate rows
        rows.append((cur, row))
ie, isamp, ite = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[ie]) for _, r in rows); tots = sum(f(r[isamp]) for _, r in rows)
print(f"{tot:.0f} warp instructions, {tots:.0f} samples")
for fn, r in rows:
    inst = f(r[ie])
    if 100 * inst / max(tot, 1) >= thr or 100 * f(r[isamp]) / max(tots, 1) >= thr:
        print(f"{fn:12s}{int(r[0]):5d} inst {100*inst/max(tot,1):5.1f}% samples {100*f(r[isamp])/max(tots,1):5.1f}% lanes {f(r[ite])/max(inst,1):4.1f} | {r[1].strip()[:130]}")
