"""oracle/oracle.py -- ctypes wrapper around oracle/liboracle.so (the CPU checker) and around the real
reference binary oracle/_ref/bcalm.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under bcalm_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
import time
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "bcalm")


class _Link(C.Structure):
    _fields_ = [("frm", C.c_uint32), ("to", C.c_uint32), ("from_sign", C.c_uint8), ("to_sign", C.c_uint8),
                ("pad", C.c_uint8 * 2)]


class _Result(C.Structure):
    _fields_ = [
        ("k", C.c_int),
        ("n_bases", C.c_uint64), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64), ("n_solid", C.c_uint64),
        ("solid_kmers", C.POINTER(C.c_uint64)), ("solid_counts", C.POINTER(C.c_uint32)),
        ("n_unitigs", C.c_uint64), ("total_nt", C.c_uint64),
        ("seq", C.POINTER(C.c_char)), ("seq_off", C.POINTER(C.c_uint64)), ("kc", C.POINTER(C.c_uint64)),
        ("abund", C.POINTER(C.c_uint32)), ("circular", C.POINTER(C.c_uint8)),
        ("n_links", C.c_uint64), ("links", C.POINTER(_Link)),
    ]


class _ResultW(C.Structure):
    _fields_ = [
        ("k", C.c_int),
        ("n_bases", C.c_uint64), ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64), ("n_solid", C.c_uint64),
        ("solid_kmers", C.POINTER(C.c_uint64)), ("solid_counts", C.POINTER(C.c_uint32)),
        ("n_unitigs", C.c_uint64), ("total_nt", C.c_uint64),
        ("seq", C.POINTER(C.c_char)), ("seq_off", C.POINTER(C.c_uint64)), ("kc", C.POINTER(C.c_uint64)),
        ("abund", C.POINTER(C.c_uint32)),
        ("n_links", C.c_uint64), ("links", C.POINTER(_Link)),
    ]


LINK_DTYPE = np.dtype([("from", "<u4"), ("to", "<u4"), ("from_sign", "u1"), ("to_sign", "u1"), ("pad", "u1", (2,))])


def build(force: bool = False) -> str:
    """Compile oracle/cdbg_oracle.c with gcc (seconds)."""
    srcs = [os.path.join(HERE, "cdbg_oracle.c"), os.path.join(HERE, "cdbg_oracle_wide.c")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB_PATH, *srcs])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_Result)]
        L.orc_build.restype = C.c_int
        L.orc_free.argtypes = [C.POINTER(_Result)]
        L.orc_extract.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64]
        L.orc_extract.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_minimizer.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_minimizer.restype = C.c_uint32
        L.orc_mmer_rank.argtypes = [C.c_uint32]
        L.orc_mmer_rank.restype = C.c_uint32
        L.orc_km_string.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_int]
        L.orcw_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_ResultW)]
        L.orcw_build.restype = C.c_int
        L.orcw_free.argtypes = [C.POINTER(_ResultW)]
        _lib = L
    return _lib


@dataclass
class Graph:
    """A compacted de Bruijn graph in array form (what both the oracle and the product return)."""
    k: int
    seq: np.ndarray          # uint8 ASCII, concatenated
    seq_off: np.ndarray      # uint64 [U+1]
    kc: np.ndarray           # uint64 [U]
    abund: np.ndarray | None  # uint32 per k-mer, unitig order (or None)
    links: np.ndarray        # LINK_DTYPE
    stats: dict = field(default_factory=dict)
    solid_kmers: np.ndarray | None = None
    solid_counts: np.ndarray | None = None

    @property
    def n_unitigs(self) -> int:
        return len(self.seq_off) - 1


def _as_bytes(reads) -> np.ndarray:
    if isinstance(reads, np.ndarray):
        return np.ascontiguousarray(reads, dtype=np.uint8)
    if isinstance(reads, (bytes, bytearray)):
        return np.frombuffer(bytes(reads), dtype=np.uint8)
    if isinstance(reads, str):
        return np.frombuffer(reads.encode(), dtype=np.uint8)
    # list of strings -> newline separated
    return np.frombuffer(("\n".join(reads) + "\n").encode(), dtype=np.uint8)


def run(reads, k: int, abundance_min: int = 2, abundance_max: int = 2**31 - 1, stages: int = 7) -> Graph:
    """Run the C oracle.  `reads`: separator-delimited byte stream (bytes / np.uint8) or a list of strings."""
    b = _as_bytes(reads)
    res = _Result()
    rc = lib().orc_build(b.ctypes.data, b.size, k, abundance_min, abundance_max, stages, C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_build failed rc={rc}")
    try:
        U, S = res.n_unitigs, res.n_solid
        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()
        g = Graph(
            k=k,
            seq=arr(res.seq, res.total_nt, np.uint8) if stages & 2 else np.zeros(0, np.uint8),
            seq_off=arr(res.seq_off, U + 1, np.uint64) if stages & 2 else np.zeros(1, np.uint64),
            kc=arr(res.kc, U, np.uint64),
            abund=arr(res.abund, S, np.uint32) if stages & 2 else None,
            links=arr(res.links, res.n_links, LINK_DTYPE),
            stats=dict(n_bases=res.n_bases, n_instances=res.n_instances, n_distinct=res.n_distinct, n_solid=S,
                       n_unitigs=U, n_links=res.n_links, total_nt=res.total_nt),
            solid_kmers=arr(res.solid_kmers, S, np.uint64),
            solid_counts=arr(res.solid_counts, S, np.uint32),
        )
    finally:
        lib().orc_free(C.byref(res))
    return g


def run_wide(reads, k: int, abundance_min: int = 2, abundance_max: int = 2**31 - 1, stages: int = 7) -> Graph:
    """The multi-word oracle (cdbg_oracle_wide.c), any k <= 127.  solid_kmers has shape (S, 4): little-endian 64-bit words."""
    b = _as_bytes(reads)
    res = _ResultW()
    rc = lib().orcw_build(b.ctypes.data, b.size, k, abundance_min, abundance_max, stages, C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orcw_build failed rc={rc}")
    try:
        U, S = res.n_unitigs, res.n_solid
        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()
        g = Graph(
            k=k,
            seq=arr(res.seq, res.total_nt, np.uint8) if stages & 2 else np.zeros(0, np.uint8),
            seq_off=arr(res.seq_off, U + 1, np.uint64) if stages & 2 else np.zeros(1, np.uint64),
            kc=arr(res.kc, U, np.uint64),
            abund=arr(res.abund, S, np.uint32) if stages & 2 else None,
            links=arr(res.links, res.n_links, LINK_DTYPE),
            stats=dict(n_bases=res.n_bases, n_instances=res.n_instances, n_distinct=res.n_distinct, n_solid=S,
                       n_unitigs=U, n_links=res.n_links, total_nt=res.total_nt),
            solid_kmers=arr(res.solid_kmers, S * 4, np.uint64).reshape(-1, 4),
            solid_counts=arr(res.solid_counts, S, np.uint32),
        )
    finally:
        lib().orcw_free(C.byref(res))
    return g


def run_any(reads, k: int, abundance_min: int = 2, **kw) -> Graph:
    """Narrow oracle for k <= 31, wide oracle above."""
    return run(reads, k, abundance_min, **kw) if k <= 31 else run_wide(reads, k, abundance_min, **kw)


def extract(reads, k: int) -> np.ndarray:
    """All valid canonical k-mer occurrences, in input order."""
    b = _as_bytes(reads)
    out = np.empty(max(1, b.size), dtype=np.uint64)
    n = lib().orc_extract(b.ctypes.data, b.size, k, out.ctypes.data, out.size)
    return out[:n].copy()


def minimizer(x: int, n: int, m: int) -> int:
    return int(lib().orc_minimizer(int(x), n, m))


def revcomp(x: int, k: int) -> int:
    return int(lib().orc_revcomp(int(x), k))


# --------------------------------------------------------------------------- the real reference binary

def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def write_fasta(path: str, reads) -> None:
    """reads: list of strings or separator-delimited bytes."""
    if isinstance(reads, (list, tuple)):
        with open(path, "w") as f:
            for r in reads:
                f.write(">r\n" + r + "\n")
        return
    b = _as_bytes(reads).tobytes()
    with open(path, "wb") as f:
        for line in b.split(b"\n"):
            if line:
                f.write(b">r\n" + line + b"\n")


def run_ref(fasta_path: str, k: int, abundance_min: int, nb_cores: int = 0, extra=(), keep_dir: str | None = None,
            timeout: float | None = None):
    """Run oracle/_ref/bcalm on a FASTA file in a scratch CWD.  Returns (unitigs.fa text, wall seconds, log)."""
    if not have_ref():
        raise FileNotFoundError(REF_BIN)
    d = keep_dir or tempfile.mkdtemp(prefix="bcalm_ref_")
    try:
        cmd = [REF_BIN, "-in", os.path.abspath(fasta_path), "-kmer-size", str(k), "-abundance-min", str(abundance_min),
               "-out", os.path.join(d, "ref"), "-out-tmp", d, "-out-dir", d, "-nb-cores", str(nb_cores), *extra]
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=timeout)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError(f"reference bcalm failed rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
        with open(os.path.join(d, "ref.unitigs.fa")) as f:
            text = f.read()
        return text, dt, p.stdout
    finally:
        if keep_dir is None:
            shutil.rmtree(d, ignore_errors=True)
