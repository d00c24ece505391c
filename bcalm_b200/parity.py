"""Parity relation between two compacted de Bruijn graphs.

The reference's own output is not byte-deterministic between two runs (unitig ids, strand, order of ``L:`` tokens and the
cut point of circular unitigs depend on thread scheduling: G/bcalm2/bglue_algo.hpp:73-90, bglue_algo.cpp:400-416,
G/debruijn/impl/LinkTigs.cpp:120-121).  What IS invariant, and what "identical to the reference" means here:

  1. the multiset of ``(canon(seq), LN, KC, km)`` with ``canon(seq) = min(seq, revcomp(seq))``; a circular unitig (one whose
     only END link is ``L:+:self:+`` and only BEGIN link is ``L:-:self:-``) is additionally minimised over the rotations
     of its k-mer cycle;
  2. the multiset of links re-expressed on canonical sequences, a sign being flipped when that node's canonical form is
     the reverse complement of what was printed.

This module parses ``unitigs.fa`` text (the reference's output grammar), converts array-form graphs to the same
records, and computes that canonical form.  It imports nothing from ``oracle/``.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def revcomp_str(s: str) -> str:
    return s.encode().translate(_COMP)[::-1].decode()


def km_string(kc: int, n_kmers: int, abund=None) -> str:
    """``km:f:`` exactly as the reference prints it: float32 running sum / float32 n, ``%.1f`` of the promoted float
    (G/bcalm2/bglue_algo.cpp:131-145 and :75-81).  With `abund` the float32 adds are replayed sequentially; without it
    the sum is taken as float32(kc), which is identical while kc < 2**24."""
    if abund is not None and kc >= (1 << 24):
        s = np.float32(0)
        for a in abund:
            s = np.float32(s + np.float32(int(a)))
    else:
        s = np.float32(kc)
    return "%.1f" % float(np.float32(s / np.float32(n_kmers)))


def parse_unitigs_fa(text: str):
    """-> list of records dict(id, seq, ln, kc, km, links=[(from_sign, to_id, to_sign)], ab=[...] or None)."""
    recs = []
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        h = lines[i]
        if not h.startswith(">"):
            i += 1
            continue
        seq = lines[i + 1].strip() if i + 1 < len(lines) else ""
        i += 2
        toks = h[1:].split()
        r = dict(id=int(toks[0]), seq=seq, ln=None, kc=None, km=None, links=[], ab=None)
        j = 1
        while j < len(toks):
            t = toks[j]
            if t.startswith("LN:i:"):
                r["ln"] = int(t[5:])
            elif t.startswith("KC:i:"):
                r["kc"] = int(t[5:])
            elif t.startswith("km:f:"):
                r["km"] = t[5:]
            elif t.startswith("ab:Z:"):
                ab = [int(t[5:])] if t[5:] else []
                while j + 1 < len(toks) and toks[j + 1].isdigit():
                    j += 1
                    ab.append(int(toks[j]))
                r["ab"] = ab
            elif t.startswith("L:"):
                _, s1, to, s2 = t.split(":")
                r["links"].append((s1, int(to), s2))
            j += 1
        recs.append(r)
    return recs


def abundance_lists(recs):
    """Records with `ab` lists -> ({canonical unitig string: abundance list in that orientation}, sorted multiset of the sorted lists).
    A circular unitig may be spelled from any cut point, so only the second part is comparable for it; compare the first on the
    strings two outputs have in common."""
    by_seq, sorted_lists = {}, []
    for r in recs:
        s, a = r["seq"], list(r["ab"])
        rc = revcomp_str(s)
        if rc < s:
            s, a = rc, a[::-1]
        elif rc == s:
            a = min(a, a[::-1])
        by_seq[s] = a
        sorted_lists.append(tuple(sorted(a)))
    return by_seq, sorted(sorted_lists)


def diff_abundance_lists(mine, theirs) -> str:
    """'' when two record lists (graph_to_records(..., with_ab=True)) carry the same per-k-mer abundance lists."""
    a, b = abundance_lists(mine), abundance_lists(theirs)
    if a[1] != b[1]:
        return "multisets of abundance lists differ"
    for s in set(a[0]) & set(b[0]):
        if a[0][s] != b[0][s]:
            return f"abundance list of {s[:60]}... differs: {a[0][s][:20]} vs {b[0][s][:20]}"
    return ""


def graph_to_records(g, with_ab: bool = False):
    """Array-form graph (fields k, seq, seq_off, kc, abund, links) -> the same record list as parse_unitigs_fa."""
    k = g.k
    seq = bytes(np.asarray(g.seq, dtype=np.uint8)).decode()
    off = np.asarray(g.seq_off, dtype=np.int64)
    U = len(off) - 1
    recs = []
    links_of = [[] for _ in range(U)]
    L = g.links
    if L is not None and len(L):
        fr = np.asarray(L["from"]); to = np.asarray(L["to"]); s1 = np.asarray(L["from_sign"]); s2 = np.asarray(L["to_sign"])
        for a, b, c, d in zip(fr.tolist(), to.tolist(), s1.tolist(), s2.tolist()):
            links_of[a].append((chr(c), b, chr(d)))
    kc = np.asarray(g.kc).tolist()
    ab_all = None if g.abund is None else np.asarray(g.abund)
    for u in range(U):
        s = seq[off[u]:off[u + 1]]
        n = len(s) - k + 1
        ab = None
        if ab_all is not None:
            a0 = int(off[u]) - u * (k - 1)
            ab = ab_all[a0:a0 + n]
        recs.append(dict(id=u, seq=s, ln=len(s), kc=int(kc[u]), km=km_string(int(kc[u]), n, ab), links=links_of[u],
                         ab=(ab.tolist() if (with_ab and ab is not None) else None)))
    return recs


def _is_circular(r) -> bool:
    b = [l for l in r["links"] if l[0] == "-"]
    e = [l for l in r["links"] if l[0] == "+"]
    return len(b) == 1 and len(e) == 1 and e[0] == ("+", r["id"], "+") and b[0] == ("-", r["id"], "-")


def _min_rotation(cyc: str) -> str:
    """Lexicographically least rotation (Booth would be O(n); cycles are short and rare so use doubling + min)."""
    n = len(cyc)
    d = cyc + cyc
    best = min(range(n), key=lambda i: d[i:i + n])
    return d[best:best + n]


def canonical_form(recs, k: int, check_ids: bool = True):
    """-> (Counter of nodes (canon_seq, ln, kc, km), Counter of links (canon_from, sign, canon_to, sign))."""
    if check_ids:
        ids = sorted(r["id"] for r in recs)
        assert ids == list(range(len(recs))), "unitig ids must be exactly 0..U-1"
    canon = {}
    flipped = {}
    palin = {}
    nodes = Counter()
    for r in recs:
        s = r["seq"].upper()
        assert r["ln"] is None or r["ln"] == len(s), f"LN mismatch for unitig {r['id']}"
        if _is_circular(r):
            cyc = s[: len(s) - (k - 1)]
            a = _min_rotation(cyc)
            b = _min_rotation(revcomp_str(s)[: len(s) - (k - 1)])
            c = min(a, b)
            c = c + c[: k - 1] if len(c) >= k - 1 else (c * k)[: len(c) + k - 1]
            canon[r["id"]] = c
            flipped[r["id"]] = False
            palin[r["id"]] = True  # both self links exist in both orientations: signs are normalised
        else:
            rc = revcomp_str(s)
            c = min(s, rc)
            canon[r["id"]] = c
            flipped[r["id"]] = c != s
            palin[r["id"]] = s == rc
        nodes[(canon[r["id"]], len(s), r["kc"], r["km"])] += 1
    links = Counter()

    def fix(sign, u):
        if palin[u]:
            return "+"
        if flipped[u]:
            return "-" if sign == "+" else "+"
        return sign

    for r in recs:
        u = r["id"]
        circ = _is_circular(r)
        for (s1, v, s2) in r["links"]:
            if circ:
                links[(canon[u], "o", canon[v], "o")] += 1
            else:
                links[(canon[u], fix(s1, u), canon[v], fix(s2, v))] += 1
    return nodes, links


def diff(recs_a, recs_b, k: int, max_show: int = 5) -> str:
    """'' when the two graphs are equal under the parity relation, else a human-readable description."""
    na, la = canonical_form(recs_a, k)
    nb, lb = canonical_form(recs_b, k)
    msgs = []
    if na != nb:
        only_a = list((na - nb).elements())[:max_show]
        only_b = list((nb - na).elements())[:max_show]
        msgs.append(f"node multisets differ: |A|={sum(na.values())} |B|={sum(nb.values())}\n  only in A: {only_a}\n  only in B: {only_b}")
    if la != lb:
        only_a = list((la - lb).elements())[:max_show]
        only_b = list((lb - la).elements())[:max_show]
        msgs.append(f"link multisets differ: |A|={sum(la.values())} |B|={sum(lb.values())}\n  only in A: {only_a}\n  only in B: {only_b}")
    return "\n".join(msgs)


def format_unitigs_fa(recs, all_abundance_counts: bool = False) -> str:
    """Write records in the reference's header grammar."""
    out = []
    for r in recs:
        if all_abundance_counts and r.get("ab") is not None:
            h = f">{r['id']} LN:i:{len(r['seq'])} ab:Z:" + " ".join(str(a) for a in r["ab"])
        else:
            h = f">{r['id']} LN:i:{len(r['seq'])} KC:i:{r['kc']} km:f:{r['km']}"
        if r["links"]:
            h += " " + " ".join(f"L:{a}:{b}:{c}" for a, b, c in r["links"])
        out.append(h)
        out.append(r["seq"])
    return "\n".join(out) + ("\n" if out else "")
