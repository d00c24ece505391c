#!/usr/bin/env python
"""Per-kernel SASS summary of the built library: instruction count, the mnemonics that matter for the design claims (TMA bulk copies
UBLKCP, mbarrier SYNCS, shared-memory atomics ATOMS / ATOMS.CAS, global atomics ATOMG / RED, shuffles, votes, PDL) and the first
lines of evidence for each.      python scripts/sass_excerpt.py > profiles/rNN_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "bcalm_b200", "libcdbg_b200.so")
want = sys.argv[1:] or ["sk1_extract", "sk2p_hist", "sk2p_scatter", "sk2c_count", "sk2d_count", "k3_ends", "k5p_hist", "k5p_scatter", "k5c_join", "k5cs_join", "k6_local",
                        "k6_jump_list", "k6_resolve", "k6_jump", "k6_finalize", "k8_spell", "k9_translate", "k9_group", "k1t_single", "k2c_count", "parse_emit",
                        "kw1_extract", "kw2_insert", "glue_assemble", "sk_compact"]
KEY = ["UBLKCP", "SYNCS", "ELECT", "ACQBULK", "PREEXIT", "ATOMS.CAS", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "MATCH", "REDUX", "LDS", "STS", "LDG", "STG", "BAR", "IMAD", "LOP3", "SHF", "BREV", "POPC", "FLO"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a); one block per kernel instantiation")
cur, body = None, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); body[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        body[cur].append(line)
dem = subprocess.run(["c++filt"], input="\n".join(body), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(body, dem))
for fn, lines in body.items():
    d = names[fn]
    short = d.split("(")[0].replace("void ", "")
    if not any(short.startswith(w) for w in want):
        continue
    ops = collections.Counter()
    first = {}
    for l in lines:
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1)
        for k in KEY:
            if op == k or op.startswith(k + "."):
                ops[k] += 1
                first.setdefault(k, l.strip()[:150])
                break
    print(f"\n== {short}   [{len(lines)} instructions]")
    print("   " + "  ".join(f"{k}:{ops[k]}" for k in KEY if ops[k]))
    for k in ("UBLKCP", "SYNCS", "ATOMS.CAS", "ATOMS", "RED", "ATOMG", "PREEXIT", "ACQBULK"):
        if k in first:
            print(f"   {k:10s} {first[k]}")
