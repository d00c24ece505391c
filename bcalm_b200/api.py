"""Python host mirror of the C ABI (include/cdbg.h) via ctypes.

`Cdbg` mirrors the reference's stage sequence for this path -- SortingCountAlgorithm -> bcalm2 -> bglue -> link_tigs
(G/debruijn/impl/UnitigsConstructionAlgorithm.cpp:92-122) -- as count() / compact() / link() / unitigs(); `build_unitigs`
is the one-call form the `bcalm` CLI uses.  There is no CPU fallback: if the CUDA library or a device is missing these
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcdbg_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
FLAG_ALL_ABUNDANCE_COUNTS, FLAG_VERBOSE = 1, 2


class CdbgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cdbg error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("abundance_min", C.c_int32), ("abundance_max", C.c_int32),
                ("minimizer_size", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32), ("rank", C.c_int32),
                ("world_size", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_bytes", "n_instances", "n_distinct", "n_solid", "n_junction_records",
                                           "n_compactions", "n_cycles", "n_unitigs", "total_nt", "n_links")] + \
               [(n, C.c_uint32) for n in ("n_count_bins", "n_join_bins", "rank_rounds", "kernel_launches")] + \
               [(n, C.c_float) for n in ("ms_push", "ms_count", "ms_compact", "ms_link", "ms_fetch", "ms_k1_hist",
                                         "ms_k1_scatter", "ms_k2", "ms_k3", "ms_k5", "ms_k6_rank", "ms_k6_spell",
                                         "ms_k2_hist", "ms_k2_scatter", "ms_k2_table")] + \
               [(n, C.c_uint32) for n in ("n_k2_splits", "k1_fallback", "zero_copy", "n_k5_splits", "k2p_fallback", "sk_m")] + \
               [("n_superkmers", C.c_uint64)] + [(n, C.c_float) for n in ("ms_x1", "ms_x2", "ms_x3", "ms_glue", "ms_xlink")] + [("n_pieces", C.c_uint32), ("sk_waves", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class _Link(C.Structure):
    _fields_ = [("frm", C.c_uint32), ("to", C.c_uint32), ("from_sign", C.c_uint8), ("to_sign", C.c_uint8), ("pad", C.c_uint8 * 2)]


class UnitigView(C.Structure):
    _fields_ = [("kmer_size", C.c_int32), ("n_unitigs", C.c_uint64), ("total_nt", C.c_uint64), ("seq", C.c_void_p),
                ("seq_off", C.c_void_p), ("kc", C.c_void_p), ("abund", C.c_void_p), ("n_links", C.c_uint64), ("links", C.c_void_p), ("link_off", C.c_void_p),
                ("first_unitig", C.c_uint64), ("n_unitigs_total", C.c_uint64)]


LINK_DTYPE = np.dtype([("from", "<u4"), ("to", "<u4"), ("from_sign", "u1"), ("to_sign", "u1"), ("pad", "u1", (2,))])

EXPORTS = ["cdbg_create", "cdbg_destroy", "cdbg_reset", "cdbg_last_error", "cdbg_push_reads", "cdbg_reserve_reads", "cdbg_push_sequence_text", "cdbg_get_pushed_reads", "cdbg_count", "cdbg_compact",
           "cdbg_link", "cdbg_get_unitigs", "cdbg_get_stats", "cdbg_build", "cdbg_get_solid_kmers", "cdbg_set_solid_kmers", "cdbg_synth_reads",
           "cdbg_copy_to_host", "cdbg_host_alloc", "cdbg_host_free", "cdbg_version", "cdbg_nccl_unique_id", "cdbg_shard_init", "cdbg_synth_reads_range", "cdbg_owner_of", "cdbg_plan_shard_exchange"]

_lib = None


def load_library():
    """Load libcdbg_b200.so (built by `python -m bcalm_b200.build` / __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CdbgError(-2, f"{LIB_PATH} is missing: build it with `python -m bcalm_b200.build` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.cdbg_create.argtypes = [P(C.c_void_p), P(Params)]
    L.cdbg_destroy.argtypes = [C.c_void_p]
    L.cdbg_destroy.restype = None
    L.cdbg_reset.argtypes = [C.c_void_p]
    L.cdbg_last_error.argtypes = [C.c_void_p]
    L.cdbg_last_error.restype = C.c_char_p
    L.cdbg_push_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.cdbg_reserve_reads.argtypes = [C.c_void_p, C.c_uint64]
    L.cdbg_get_pushed_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, P(C.c_uint64)]
    L.cdbg_push_sequence_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, P(C.c_uint64)]
    for n in ("cdbg_count", "cdbg_compact", "cdbg_link", "cdbg_get_stats"):
        getattr(L, n).argtypes = [C.c_void_p, P(Stats)]
    L.cdbg_get_unitigs.argtypes = [C.c_void_p, P(UnitigView)]
    L.cdbg_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, P(UnitigView), P(Stats)]
    L.cdbg_get_solid_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, P(C.c_uint64)]
    L.cdbg_set_solid_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.cdbg_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, P(C.c_void_p), P(C.c_uint64)]
    L.cdbg_synth_reads_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, P(C.c_void_p), P(C.c_uint64)]
    L.cdbg_owner_of.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_uint32]
    L.cdbg_owner_of.restype = C.c_uint32
    L.cdbg_plan_shard_exchange.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cdbg_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.cdbg_host_alloc.argtypes = [C.c_uint64]
    L.cdbg_host_alloc.restype = C.c_void_p
    L.cdbg_host_free.argtypes = [C.c_void_p]
    L.cdbg_host_free.restype = None
    L.cdbg_version.restype = C.c_char_p
    L.cdbg_nccl_unique_id.argtypes = [C.c_void_p]
    L.cdbg_shard_init.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


@dataclass
class Graph:
    """Array form of <prefix>.unitigs.fa (same fields as the oracle's Graph, so parity.graph_to_records takes either)."""
    k: int
    seq: np.ndarray
    seq_off: np.ndarray
    kc: np.ndarray
    abund: np.ndarray | None
    links: np.ndarray
    stats: dict = field(default_factory=dict)
    link_off: np.ndarray | None = None
    first_unitig: int = 0          # sharded mode: global id of local unitig 0 (links always carry global ids)
    n_unitigs_total: int = 0

    @property
    def n_unitigs(self) -> int:
        return len(self.seq_off) - 1


def _as_u8(reads) -> np.ndarray:
    if isinstance(reads, np.ndarray):
        return np.ascontiguousarray(reads, dtype=np.uint8)
    if isinstance(reads, (bytes, bytearray, memoryview)):
        return np.frombuffer(reads, dtype=np.uint8)
    if isinstance(reads, str):
        return np.frombuffer(reads.encode(), dtype=np.uint8)
    return np.frombuffer(("\n".join(reads) + "\n").encode(), dtype=np.uint8)


class Cdbg:
    """One context = one GPU.  Usage: g = Cdbg(k=31, abundance_min=2).build(reads)"""

    def __init__(self, k: int = 31, abundance_min: int = 2, abundance_max: int = 0, minimizer_size: int = 0, device: int = 0,
                 all_abundance_counts: bool = False, verbose: bool = False, rank: int = 0, world_size: int = 1):
        self.lib = load_library()
        self.k = k
        flags = (FLAG_ALL_ABUNDANCE_COUNTS if all_abundance_counts else 0) | (FLAG_VERBOSE if verbose else 0)
        self.params = Params(k, abundance_min, abundance_max, minimizer_size, device, flags, rank, world_size)
        self.rank, self.world_size = rank, world_size
        self.h = C.c_void_p()
        rc = self.lib.cdbg_create(C.byref(self.h), C.byref(self.params))
        if rc != 0:
            msg = self.lib.cdbg_last_error(self.h).decode() if self.h else "cdbg_create failed"
            if self.h:
                self.lib.cdbg_destroy(self.h)
                self.h = C.c_void_p()
            raise CdbgError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cdbg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise CdbgError(rc, self.lib.cdbg_last_error(self.h).decode())

    # ---- sharded mode (one Cdbg per GPU / process)
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = load_library().cdbg_nccl_unique_id(buf)
        if rc != 0:
            raise CdbgError(rc, "ncclGetUniqueId failed")
        return bytes(buf)

    def shard_init(self, unique_id: bytes):
        """Collective: every rank passes the 128-byte id obtained by rank 0 from nccl_unique_id()."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.cdbg_shard_init(self.h, buf))

    def shard_init_torch(self):
        """Bootstrap through an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist
        obj = [self.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        self.shard_init(obj[0])

    # ---- stage calls
    def reset(self):
        self._ck(self.lib.cdbg_reset(self.h))

    def push_reads(self, reads):
        b = _as_u8(reads)
        self._keep = b
        self._ck(self.lib.cdbg_push_reads(self.h, b.ctypes.data, b.size, MEM_HOST))

    def push_sequence_text(self, text) -> int:
        """Raw FASTA / FASTQ text (bytes of the file) parsed on the GPU; returns the number of records."""
        b = _as_u8(text)
        self._keep = b
        n = C.c_uint64()
        self._ck(self.lib.cdbg_push_sequence_text(self.h, b.ctypes.data, b.size, MEM_HOST, C.byref(n)))
        return int(n.value)

    def pushed_reads(self) -> bytes:
        n = C.c_uint64()
        self._ck(self.lib.cdbg_get_pushed_reads(self.h, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._ck(self.lib.cdbg_get_pushed_reads(self.h, buf.ctypes.data, n.value, C.byref(n)))
        return buf.tobytes()

    def reserve_reads(self, n_bytes: int):
        """Make room on the device for n_bytes of reads before a series of pushes."""
        self._ck(self.lib.cdbg_reserve_reads(self.h, int(n_bytes)))

    def push_device(self, dptr: int, n_bytes: int):
        self._ck(self.lib.cdbg_push_reads(self.h, C.c_void_p(dptr), n_bytes, MEM_DEVICE))

    def count(self) -> dict:
        s = Stats()
        self._ck(self.lib.cdbg_count(self.h, C.byref(s)))
        return s.as_dict()

    def compact(self) -> dict:
        s = Stats()
        self._ck(self.lib.cdbg_compact(self.h, C.byref(s)))
        return s.as_dict()

    def link(self) -> dict:
        s = Stats()
        self._ck(self.lib.cdbg_link(self.h, C.byref(s)))
        return s.as_dict()

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.cdbg_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def solid_kmers(self):
        n = C.c_uint64()
        self._ck(self.lib.cdbg_get_solid_kmers(self.h, None, None, 0, C.byref(n)))
        W = 1 if self.k <= 31 else max(2, (2 * self.k + 63) // 64)
        km = np.empty(n.value * W, np.uint64)
        ct = np.empty(n.value, np.uint32)
        self._ck(self.lib.cdbg_get_solid_kmers(self.h, km.ctypes.data, ct.ctypes.data, n.value, C.byref(n)))
        return (km if W == 1 else km.reshape(-1, W)), ct

    def set_solid_kmers(self, kmers, counts):
        """Start from counted k-mers instead of reads: uint64 k-mers in either orientation (k <= 31: shape (n,); k >= 32: shape (n, W),
        little-endian words, as solid_kmers() returns them), uint32 abundances."""
        W = 1 if self.k <= 31 else max(2, (2 * self.k + 63) // 64)
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        ct = np.ascontiguousarray(counts, dtype=np.uint32)
        if ct.ndim != 1 or km.shape != ((ct.size,) if W == 1 else (ct.size, W)):
            raise ValueError(f"kmers must have shape {'(n,)' if W == 1 else f'(n, {W})'} and counts shape (n,)")
        self._ck(self.lib.cdbg_set_solid_kmers(self.h, km.ctypes.data, ct.ctypes.data, ct.size, MEM_HOST))

    def _view_to_graph(self, v: UnitigView, copy: bool = True) -> Graph:
        def arr(ptr, n, dt):
            if n == 0 or not ptr:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt)
            return a.copy() if copy else a
        U = v.n_unitigs
        S = int(v.total_nt - U * (self.k - 1)) if U else 0
        return Graph(k=self.k, seq=arr(v.seq, v.total_nt, np.uint8), seq_off=arr(v.seq_off, U + 1, np.uint64),
                     kc=arr(v.kc, U, np.uint64), abund=arr(v.abund, S, np.uint32) if v.abund else None,
                     links=arr(v.links, v.n_links, LINK_DTYPE), stats=self.stats(), link_off=arr(v.link_off, U + 1, np.uint64),
                     first_unitig=int(v.first_unitig), n_unitigs_total=int(v.n_unitigs_total))

    def unitigs(self, copy: bool = True) -> Graph:
        v = UnitigView()
        self._ck(self.lib.cdbg_get_unitigs(self.h, C.byref(v)))
        return self._view_to_graph(v, copy)

    def build(self, reads, copy: bool = True) -> Graph:
        """reads (host bytes / np.uint8 / list of str) -> Graph, through cdbg_build (H2D + all stages + D2H)."""
        b = _as_u8(reads)
        v, s = UnitigView(), Stats()
        self._ck(self.lib.cdbg_build(self.h, b.ctypes.data, b.size, MEM_HOST, C.byref(v), C.byref(s)))
        return self._view_to_graph(v, copy)

    def build_raw(self, ptr: int, n_bytes: int, location: int, fetch: bool = True):
        """Timing entry used by bench.py: no numpy conversion of the result."""
        v, s = UnitigView(), Stats()
        self._ck(self.lib.cdbg_build(self.h, C.c_void_p(ptr), n_bytes, location, C.byref(v) if fetch else None, C.byref(s)))
        return v, s

    def synth_reads(self, genome_len: int, read_len: int, n_reads: int, err_ppm: int = 10000, seed: int = 42, first_read: int = 0):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.cdbg_synth_reads_range(self.h, genome_len, read_len, first_read, n_reads, err_ppm, seed, C.byref(p), C.byref(n)))
        return p.value, n.value

    def copy_to_host(self, dptr: int, n_bytes: int) -> np.ndarray:
        out = np.empty(n_bytes, np.uint8)
        self._ck(self.lib.cdbg_copy_to_host(self.h, out.ctypes.data, C.c_void_p(dptr), n_bytes))
        return out


class PinnedBuffer:
    """Page-locked host memory from cdbg_host_alloc as a numpy uint8 array (`.array`).  Reads placed here are copied at full PCIe
    speed by push_reads and read IN PLACE by build() (no separate H2D copy, see cdbg_stats.zero_copy)."""

    def __init__(self, n_bytes: int):
        self.lib = load_library()
        self.ptr = self.lib.cdbg_host_alloc(max(1, int(n_bytes)))
        if not self.ptr:
            raise CdbgError(-3, f"cdbg_host_alloc({n_bytes}) failed")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(int(n_bytes),))

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.cdbg_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def owner_of(x: int, n_nt: int, m: int, world: int) -> int:
    """GPU owning the n_nt-mer x in a `world`-way sharded build (host-side, no GPU needed)."""
    return int(load_library().cdbg_owner_of(int(x), n_nt, m, world))


def plan_shard_exchange(world: int, rank: int, p_local: int, all_cnt: np.ndarray):
    """-> (send_cnt[world], recv_cnt[world], run_off[p_local, world], run_len[p_local, world]) for `rank`."""
    a = np.ascontiguousarray(all_cnt, dtype=np.uint64).reshape(-1)
    assert a.size == world * world * p_local
    send, recv = np.zeros(world, np.uint64), np.zeros(world, np.uint64)
    off, ln = np.zeros((p_local, world), np.uint64), np.zeros((p_local, world), np.uint64)
    rc = load_library().cdbg_plan_shard_exchange(world, rank, p_local, a.ctypes.data, send.ctypes.data, recv.ctypes.data, off.ctypes.data, ln.ctypes.data)
    if rc != 0:
        raise CdbgError(rc, "cdbg_plan_shard_exchange: bad arguments")
    return send, recv, off, ln


def merge_shards(graphs) -> Graph:
    """Concatenate the per-rank graphs of a sharded build (ordered by first_unitig) into one Graph with global ids."""
    gs = sorted(graphs, key=lambda g: (g.first_unitig, g.n_unitigs))
    k = gs[0].k
    seq = np.concatenate([g.seq for g in gs]) if gs else np.zeros(0, np.uint8)
    offs, base, expect = [np.zeros(1, np.uint64)], 0, 0
    for g in gs:
        assert g.first_unitig == expect, "shards do not tile the unitig id range"
        offs.append(g.seq_off[1:].astype(np.uint64) + np.uint64(base))
        base += len(g.seq)
        expect += g.n_unitigs
    links = np.concatenate([g.links for g in gs])
    order = np.argsort(links["from"], kind="stable")
    links = links[order]
    link_off = np.zeros(expect + 1, np.uint64)
    np.cumsum(np.bincount(links["from"], minlength=expect), out=link_off[1:])
    return Graph(k=k, seq=seq, seq_off=np.concatenate(offs), kc=np.concatenate([g.kc for g in gs]),
                 abund=np.concatenate([g.abund for g in gs]) if gs and all(g.abund is not None for g in gs) else None, links=links,
                 stats={"shards": [g.stats for g in gs]}, link_off=link_off, first_unitig=0, n_unitigs_total=expect)


def build_unitigs(reads, k: int = 31, abundance_min: int = 2, **kw) -> Graph:
    with Cdbg(k=k, abundance_min=abundance_min, **kw) as c:
        return c.build(reads)
