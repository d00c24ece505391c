# dev helper: per-source-line instruction counts and stall samples from an ncu report (needs -lineinfo + --import-source on)
#   python scripts/ncu_src.py report.ncu-rep [top]
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows, hdr, cur_file = [], None, ""
for row in csv.reader(io.StringIO(txt)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr and len(row) == len(hdr) and row[0] not in ("", "Line No"):
        rows.append((cur_file, row))
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ite = hdr.index("Thread Instructions Executed")
def f(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(f(r[ie]) for _, r in rows); tots = sum(f(r[isamp]) for _, r in rows)
print(f"{tot:.0f} warp instructions, {tots:.0f} samples")
best = sorted(rows, key=lambda fr: -f(fr[1][isamp]))[:top]
for fn, r in sorted(best, key=lambda fr: (fr[0], int(fr[1][0]))):
    inst = f(r[ie]); thr = f(r[ite])
    print(f"{fn:12s}{int(r[0]):5d} inst {100*inst/max(tot,1):5.1f}% samples {100*f(r[isamp])/max(tots,1):5.1f}% lanes {thr/max(inst,1):4.1f} | {r[1].strip()[:140]}")
