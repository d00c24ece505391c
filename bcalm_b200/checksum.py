"""Order-, id-, strand- and cut-point-independent checksum of a compacted de Bruijn graph (k <= 31), computed with torch.

Test / bench infrastructure (the "checksum of checksums" of the parity plan): two graphs that are equal under the parity
relation of `parity.py` have equal checksums, at any size -- `parity.diff` walks Python records and is only
practical up to ~10^5 unitigs, this runs on the GPU (or on CPU tensors in the CPU suite) over 10^8..10^9 nucleotides.

  nodes : every unitig contributes mix(sum over its k-mers of mix(canonical k-mer value), LN, KC).  The k-mer sum does not depend
          on the strand the unitig is written in nor, for a circular unitig, on where the cycle was cut.
  links : a link (u, su, v, sv) is described by the two ORIENTED k-mers that meet at the junction -- a = last k-mer of u read in
          orientation su, b = first k-mer of v read in orientation sv -- which does not depend on unitig ids or on the strand a
          unitig happens to be written in.  The two self links of an isolated cycle (`L:+:u:+`, `L:-:u:-`) depend on the cut point
          and contribute a constant.  `bad_overlaps` counts links whose two k-mers do not overlap by k-1 (must be 0).
k-mer values use the reference's encoding (A=0 C=1 T=2 G=3, first nucleotide most significant; G/tools/misc/api/Data.hpp:185).
"""
from __future__ import annotations

import numpy as np

_M1 = -49064778989728563          # 0xff51afd7ed558ccd as int64
_M2 = -4265267296055464877        # 0xc4ceb9fe1a85ec53 as int64
_G = 0x9E3779B97F4A7C15 - (1 << 64)   # golden-ratio constant as int64
_SELF = 0x5bd1e9955bd1e995


def _mix(x):
    """murmur3 fmix64 on int64 tensors (logical shifts emulated; multiplication wraps)."""
    m31 = (1 << 31) - 1
    x = x ^ ((x >> 33) & m31)
    x = x * _M1
    x = x ^ ((x >> 33) & m31)
    x = x * _M2
    x = x ^ ((x >> 33) & m31)
    return x


def _windows(code, k, torch):
    """fwd[i], rc[i] = value of the k-mer starting at i and of its reverse complement, for every i (garbage past N - k)."""
    N = code.numel()
    pad = torch.zeros(64, dtype=torch.int64, device=code.device)
    f = {1: torch.cat([code, pad])}
    r = {1: torch.cat([code ^ 2, pad])}
    p = 1
    while 2 * p <= k:
        a, b = f[p], r[p]
        f[2 * p] = torch.cat([(a[:-p] << (2 * p)) | a[p:], pad[:p]])
        r[2 * p] = torch.cat([(b[p:] << (2 * p)) | b[:-p], pad[:p]])
        p *= 2
    fwd = rc = None
    off = 0
    while p >= 1:
        if k & p:
            a, b = f[p][off:off + N], r[p][off:off + N]
            if a.numel() < N:
                a = torch.cat([a, pad[:N - a.numel()]]); b = torch.cat([b, pad[:N - b.numel()]])
            fwd = a if fwd is None else ((fwd << (2 * p)) | a)
            rc = b if rc is None else ((b << (2 * off)) | rc)
            off += p
        p //= 2
    return fwd, rc


def graph_checksum(g, device="cpu", chunk_nt: int = 1 << 26) -> dict:
    """g: array-form graph (bcalm_b200.api.Graph or the oracle's): k, seq (ASCII), seq_off, kc, links (from/to/from_sign/to_sign, global ids)."""
    import torch
    k = int(g.k)
    if k > 31:
        raise ValueError("graph_checksum handles one-word k-mers (k <= 31)")
    dev = torch.device(device)
    seq_off = np.asarray(g.seq_off, dtype=np.int64)
    U = len(seq_off) - 1
    out = {"n_unitigs": int(U), "total_nt": int(seq_off[-1]) if U else 0, "sum_kc": int(np.asarray(g.kc, dtype=np.uint64).sum()) if U else 0,
           "n_links": int(len(g.links)) if g.links is not None else 0, "nodes": 0, "links": 0, "bad_overlaps": 0}
    if U == 0:
        return out
    lut = np.full(256, 0, np.int64)
    for ch, v in (("A", 0), ("C", 1), ("T", 2), ("G", 3), ("a", 0), ("c", 1), ("t", 2), ("g", 3)):
        lut[ord(ch)] = v
    seq = np.asarray(g.seq, dtype=np.uint8)
    kc = torch.from_numpy(np.asarray(g.kc, dtype=np.uint64).view(np.int64).copy()).to(dev)
    off_t = torch.from_numpy(seq_off).to(dev)
    ksum = torch.zeros(U, dtype=torch.int64, device=dev)
    F = torch.zeros(U, dtype=torch.int64, device=dev); L = torch.zeros_like(F); RF = torch.zeros_like(F); RL = torch.zeros_like(F)
    lut_t = torch.from_numpy(lut).to(dev)
    u0 = 0
    while u0 < U:                       # chunks of whole unitigs
        u1 = int(np.searchsorted(seq_off, seq_off[u0] + chunk_nt, side="right")) - 1
        u1 = min(U, max(u1, u0 + 1))
        b0, b1 = int(seq_off[u0]), int(seq_off[u1])
        code = lut_t[torch.from_numpy(seq[b0:b1].copy()).to(dev).long()]
        fwd, rc = _windows(code, k, torch)
        tok = _mix(torch.minimum(fwd, rc))
        ends = off_t[u0 + 1:u1 + 1] - b0           # exclusive end of every unitig in the chunk
        starts = off_t[u0:u1] - b0
        inval = torch.zeros(b1 - b0 + 1, dtype=torch.int64, device=dev)
        for j in range(1, k):                      # the last k-1 positions of a unitig start no k-mer
            idx = ends - j
            inval[idx[idx >= starts]] = 1
        tok = torch.where(inval[:-1] == 0, tok, torch.zeros_like(tok))
        cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(tok, 0)])
        ksum[u0:u1] = cs[ends] - cs[starts]
        F[u0:u1] = fwd[starts]; RF[u0:u1] = rc[starts]
        L[u0:u1] = fwd[ends - k]; RL[u0:u1] = rc[ends - k]
        u0 = u1
    ln = off_t[1:] - off_t[:-1]
    nh = _mix(ksum ^ _mix(ln * _G + kc * _M1))
    out["nodes"] = int(nh.sum().item()) & 0xFFFFFFFFFFFFFFFF
    if out["n_links"]:
        lk = g.links
        fr = torch.from_numpy(np.asarray(lk["from"], dtype=np.int64)).to(dev)
        to = torch.from_numpy(np.asarray(lk["to"], dtype=np.int64)).to(dev)
        s1 = torch.from_numpy(np.asarray(lk["from_sign"], dtype=np.int64)).to(dev) == ord("+")
        s2 = torch.from_numpy(np.asarray(lk["to_sign"], dtype=np.int64)).to(dev) == ord("+")
        a = torch.where(s1, L[fr], RF[fr])
        b = torch.where(s2, F[to], RL[to])
        mask = (1 << (2 * (k - 1))) - 1
        out["bad_overlaps"] = int(((a & mask) != (b >> 2)).sum().item())
        h = _mix(_mix(a) * _G + b)
        selfl = (fr == to) & (s1 == s2)
        h = torch.where(selfl, torch.full_like(h, _SELF), h)
        out["links"] = int(h.sum().item()) & 0xFFFFFFFFFFFFFFFF
    return out


def records_to_graph(recs, k: int):
    """parse_unitigs_fa records (e.g. the reference's unitigs.fa) -> an array-form graph graph_checksum accepts."""
    from types import SimpleNamespace
    from .api import LINK_DTYPE
    order = sorted(recs, key=lambda r: r["id"])
    seq = np.frombuffer("".join(r["seq"] for r in order).encode(), dtype=np.uint8)
    off = np.zeros(len(order) + 1, np.uint64)
    np.cumsum([len(r["seq"]) for r in order], out=off[1:])
    links = np.zeros(sum(len(r["links"]) for r in order), LINK_DTYPE)
    i = 0
    for r in order:
        for (s1, v, s2) in r["links"]:
            links[i] = (r["id"], v, ord(s1), ord(s2), (0, 0))
            i += 1
    return SimpleNamespace(k=k, seq=seq, seq_off=off, kc=np.array([r["kc"] for r in order], np.uint64), links=links)
