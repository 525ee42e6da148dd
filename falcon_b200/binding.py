"""ctypes binding of libfalcon_b200.so -- the counterpart of falcon_kit/falcon_kit.py.

The reference binds its C library with ``CDLL(ext_falcon.__file__)`` and attaches argtypes at
import (falcon_kit/falcon_kit.py:44-122); this module does the same for the B200 library and adds
the batched ``fcx_*`` entry points (include/falcon_b200.h).  There is no fallback: if the shared
library is missing or no CUDA device is usable, importing / creating an engine raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfalcon_b200.so")

seq_coor_t = C.c_int
base_t = C.c_uint8


# ---- struct mirrors, same layouts as falcon_kit/falcon_kit.py:19-41,86-106 -------------------
class KmerLookup(C.Structure):
    _fields_ = [("start", seq_coor_t), ("last", seq_coor_t), ("count", seq_coor_t)]


class KmerMatch(C.Structure):
    _fields_ = [("count", seq_coor_t), ("query_pos", C.POINTER(seq_coor_t)),
                ("target_pos", C.POINTER(seq_coor_t))]


class AlnRange(C.Structure):
    _fields_ = [("s1", seq_coor_t), ("e1", seq_coor_t), ("s2", seq_coor_t), ("e2", seq_coor_t),
                ("score", C.c_long)]


class ConsensusData(C.Structure):
    _fields_ = [("sequence", C.c_char_p), ("eff_cov", C.POINTER(C.c_uint))]


class Alignment(C.Structure):
    _fields_ = [("aln_str_size", seq_coor_t), ("dist", seq_coor_t), ("aln_q_s", seq_coor_t),
                ("aln_q_e", seq_coor_t), ("aln_t_s", seq_coor_t), ("aln_t_e", seq_coor_t),
                ("q_aln_str", C.c_char_p), ("t_aln_str", C.c_char_p)]


class PairInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_match", "s1", "e1", "s2", "e2", "passed_filter",
                                           "aligned", "dist", "aln_size", "q_e", "t_e", "accepted",
                                           "n_tags", "trace_cells")]


T_NAMES = ("index", "range", "dp", "traceback", "consensus", "total")
C_NAMES = ("pairs", "dp_pairs", "accepted", "trace_cells", "dp_steps", "aln_cols", "span_bases",
           "kernel_launches", "waves")


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            "falcon_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    # legacy symbols (falcon_kit/falcon_kit.py:54-122, falcon_kit/mains/consensus.py:20-23)
    lib.allocate_kmer_lookup.argtypes = [seq_coor_t]
    lib.allocate_kmer_lookup.restype = C.POINTER(KmerLookup)
    lib.init_kmer_lookup.argtypes = [C.POINTER(KmerLookup), seq_coor_t]
    lib.free_kmer_lookup.argtypes = [C.POINTER(KmerLookup)]
    lib.allocate_seq.argtypes = [seq_coor_t]
    lib.allocate_seq.restype = C.POINTER(base_t)
    lib.init_seq_array.argtypes = [C.POINTER(base_t), seq_coor_t]
    lib.free_seq_array.argtypes = [C.POINTER(base_t)]
    lib.allocate_seq_addr.argtypes = [seq_coor_t]
    lib.allocate_seq_addr.restype = C.POINTER(seq_coor_t)
    lib.free_seq_addr_array.argtypes = [C.POINTER(seq_coor_t)]
    lib.add_sequence.argtypes = [seq_coor_t, C.c_uint, C.POINTER(C.c_char), seq_coor_t,
                                 C.POINTER(seq_coor_t), C.POINTER(C.c_uint8), C.POINTER(KmerLookup)]
    lib.mask_k_mer.argtypes = [C.c_long, C.POINTER(KmerLookup), C.c_long]
    lib.find_kmer_pos_for_seq.argtypes = [C.POINTER(C.c_char), seq_coor_t, C.c_uint,
                                          C.POINTER(seq_coor_t), C.POINTER(KmerLookup)]
    lib.find_kmer_pos_for_seq.restype = C.POINTER(KmerMatch)
    lib.free_kmer_match.argtypes = [C.POINTER(KmerMatch)]
    lib.find_best_aln_range.argtypes = [C.POINTER(KmerMatch), seq_coor_t, seq_coor_t, seq_coor_t]
    lib.find_best_aln_range.restype = C.POINTER(AlnRange)
    lib.find_best_aln_range2.argtypes = [C.POINTER(KmerMatch), seq_coor_t, seq_coor_t, seq_coor_t]
    lib.find_best_aln_range2.restype = C.POINTER(AlnRange)
    lib.free_aln_range.argtypes = [C.POINTER(AlnRange)]
    lib.align.argtypes = [C.POINTER(C.c_char), C.c_long, C.POINTER(C.c_char), C.c_long, C.c_long,
                          C.c_int]
    lib.align.restype = C.POINTER(Alignment)
    lib.free_alignment.argtypes = [C.POINTER(Alignment)]
    lib.generate_consensus.argtypes = [C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_uint,
                                       C.c_double]
    lib.generate_consensus.restype = C.POINTER(ConsensusData)
    lib.free_consensus_data.argtypes = [C.POINTER(ConsensusData)]
    # batched path
    lib.fcx_version.restype = C.c_char_p
    lib.fcx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.fcx_destroy.argtypes = [C.c_void_p]
    lib.fcx_last_error.argtypes = [C.c_void_p]
    lib.fcx_last_error.restype = C.c_char_p
    lib.fcx_host_alloc.argtypes = [C.c_size_t]
    lib.fcx_host_alloc.restype = C.c_void_p
    lib.fcx_host_free.argtypes = [C.c_void_p]
    lib.fcx_pool_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.fcx_pool_reserve.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
    lib.fcx_pool_upload_part.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.fcx_pool_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
    lib.fcx_pool_commit.argtypes = [C.c_void_p]
    lib.fcx_consensus_blocks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint,
                                         C.c_uint, C.c_double, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_void_p)]
    lib.fcx_last_pair_info.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.fcx_last_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.fcx_trim_blocks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint,
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    lib.fcx_pool_truncate.argtypes = [C.c_void_p, C.c_uint32]
    lib.fcx_align_pairs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.fcx_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    lib.fcx_multi_destroy.argtypes = [C.c_void_p]
    lib.fcx_multi_last_error.argtypes = [C.c_void_p]
    lib.fcx_multi_last_error.restype = C.c_char_p
    lib.fcx_multi_device_count.argtypes = [C.c_void_p]
    lib.fcx_multi_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.fcx_multi_pool_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.fcx_multi_peer_bytes.argtypes = [C.c_void_p]
    lib.fcx_multi_peer_bytes.restype = C.c_uint64
    lib.fcx_multi_consensus_blocks.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint,
                                               C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.fcx_multi_last_pair_info.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.fcx_multi_last_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.fcx_parser_create.argtypes = [C.c_uint] * 5
    lib.fcx_parser_create.restype = C.c_void_p
    lib.fcx_parser_destroy.argtypes = [C.c_void_p]
    lib.fcx_parser_feed.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int]
    lib.fcx_parser_pending.argtypes = [C.c_void_p]
    lib.fcx_parser_stopped.argtypes = [C.c_void_p]
    lib.fcx_parser_take.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)]
    lib.fcx_dazz_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.fcx_dazz_close.argtypes = [C.c_void_p]
    lib.fcx_dazz_last_error.argtypes = [C.c_void_p]
    lib.fcx_dazz_last_error.restype = C.c_char_p
    lib.fcx_dazz_nreads.argtypes = [C.c_void_p]
    lib.fcx_dazz_nreads.restype = C.c_uint32
    lib.fcx_dazz_read_length.argtypes = [C.c_void_p, C.c_uint32]
    lib.fcx_dazz_read_length.restype = C.c_int32
    lib.fcx_dazz_upload.argtypes = [C.c_void_p, C.c_void_p]
    lib.fcx_las_open.argtypes = [C.c_void_p, C.c_char_p]
    lib.fcx_las_take.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint32, C.c_uint64,
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_int)]
    lib.fcx_pool_upload_bps.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.fcx_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.fcx_timer_start.argtypes = [C.c_void_p]
    lib.fcx_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    return lib


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


# the three names the reference exports for one and the same CDLL (falcon_kit.py:52,109,117)
def kup() -> C.CDLL:
    return lib()


DWA = kup
falcon = kup


def device_count() -> int:
    return int(lib().fcx_device_count())


class EngineError(RuntimeError):
    pass


class PinnedBuffer:
    """Page-locked host staging buffer (numpy view) for read bytes."""

    def __init__(self, nbytes: int):
        self._lib = lib()
        self.ptr = self._lib.fcx_host_alloc(max(1, nbytes))
        if not self.ptr:
            raise EngineError("fcx_host_alloc(%d) failed" % nbytes)
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, nbytes)).from_address(self.ptr))

    def close(self):
        if self.ptr:
            self._lib.fcx_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One GPU engine (one per process per GPU)."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        h = C.c_void_p()
        if self._lib.fcx_create(device, C.byref(h)) != 0:
            raise EngineError("fcx_create(%d): %s" % (device, self._lib.fcx_last_error(None).decode()))
        self._h = h
        self.n_reads = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fcx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError("%s: %s" % (what, self._lib.fcx_last_error(self._h).decode()))

    def set_option(self, name: str, value: float):
        self._check(self._lib.fcx_set_option(self._h, name.encode(), float(value)), "fcx_set_option")

    # -- pool ---------------------------------------------------------------------------
    def upload_pool_raw(self, bases_ptr: int, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        self._check(self._lib.fcx_pool_upload(self._h, bases_ptr, offsets.ctypes.data, n),
                    "fcx_pool_upload")
        self.n_reads = self._reserved = n

    def upload_pool(self, reads: Sequence[bytes]):
        offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum([len(r) for r in reads], out=offsets[1:])
        cat = b"".join(reads)
        buf = C.create_string_buffer(cat, len(cat) + 1)
        self.upload_pool_raw(C.addressof(buf), offsets)

    # -- pool assembled from parts (multi-process / multi-GPU) -------------
    def pool_reserve(self, offsets: np.ndarray) -> int:
        """Fix the pool layout from the lengths of ALL reads; returns the packed size in words."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        w = C.c_uint64()
        self._check(self._lib.fcx_pool_reserve(self._h, offsets.ctypes.data, n, C.byref(w)), "fcx_pool_reserve")
        self._reserved = n
        return int(w.value)

    def pool_upload_part(self, bases_ptr: int, offsets_part: np.ndarray, first_read: int):
        offsets_part = np.ascontiguousarray(offsets_part, dtype=np.uint64)
        self._check(self._lib.fcx_pool_upload_part(self._h, bases_ptr, offsets_part.ctypes.data, first_read,
                                                   offsets_part.shape[0] - 1), "fcx_pool_upload_part")

    def pool_device(self):
        """-> (device pointer of the packed pool, number of 32-bit words, word offsets per read)."""
        ptr, nw, wo = C.c_void_p(), C.c_uint64(), C.c_void_p()
        self._check(self._lib.fcx_pool_device(self._h, C.byref(ptr), C.byref(nw), C.byref(wo)), "fcx_pool_device")
        woff = np.ctypeslib.as_array((C.c_uint64 * (self._reserved + 1)).from_address(wo.value)).copy()
        return int(ptr.value), int(nw.value), woff

    def pool_commit(self):
        self._check(self._lib.fcx_pool_commit(self._h), "fcx_pool_commit")
        self.n_reads = self._reserved

    # -- consensus ----------------------------------------------------------------------
    def consensus_blocks_raw(self, block_off: np.ndarray, read_ids: np.ndarray, min_cov: int,
                             min_idt: float, K: int = 8, copy: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """-> (consensus bytes of all blocks back to back, n_blocks + 1 offsets).  With copy=False the
        arrays are views of the engine's own result buffers (host memory), valid until the next call."""
        block_off = np.ascontiguousarray(block_off, dtype=np.uint32)
        read_ids = np.ascontiguousarray(read_ids, dtype=np.uint32)
        nb = block_off.shape[0] - 1
        ob, oo = C.c_void_p(), C.c_void_p()
        self._check(self._lib.fcx_consensus_blocks(self._h, nb, block_off.ctypes.data,
                                                   read_ids.ctypes.data, min_cov, K, min_idt,
                                                   C.byref(ob), C.byref(oo)),
                    "fcx_consensus_blocks")
        off = np.ctypeslib.as_array((C.c_uint64 * (nb + 1)).from_address(oo.value))
        total = int(off[-1])
        if total:
            data = np.ctypeslib.as_array((C.c_uint8 * total).from_address(ob.value))
        else:
            data = np.zeros(0, dtype=np.uint8)
        return (data.copy(), off.copy()) if copy else (data, off)

    def consensus_blocks(self, blocks: Sequence[Sequence[int]], min_cov: int, min_idt: float,
                         K: int = 8) -> List[bytes]:
        block_off = np.zeros(len(blocks) + 1, dtype=np.uint32)
        np.cumsum([len(b) for b in blocks], out=block_off[1:])
        ids = np.concatenate([np.asarray(b, dtype=np.uint32) for b in blocks]) if blocks else \
            np.zeros(0, dtype=np.uint32)
        data, off = self.consensus_blocks_raw(block_off, ids, min_cov, min_idt, K)
        raw = data.tobytes()
        return [raw[int(off[i]):int(off[i + 1])] for i in range(len(blocks))]

    def trim_blocks_raw(self, block_off: np.ndarray, read_ids: np.ndarray, edge_tolerance: int = 1000,
                        trim_size: int = 50, max_n_read: int = 500, max_cov_aln: int = 0):
        """--trim on the device (get_consensus_with_trim, consensus.py:123-147): appends the trimmed
        reads to the pool and returns the new (block_off, read_ids)."""
        block_off = np.ascontiguousarray(block_off, dtype=np.uint32)
        read_ids = np.ascontiguousarray(read_ids, dtype=np.uint32)
        nb = block_off.shape[0] - 1
        ob, oi, nr = C.c_void_p(), C.c_void_p(), C.c_uint32()
        self._check(self._lib.fcx_trim_blocks(self._h, nb, block_off.ctypes.data, read_ids.ctypes.data,
                                              edge_tolerance, trim_size, max_n_read, max_cov_aln,
                                              C.byref(ob), C.byref(oi), C.byref(nr)), "fcx_trim_blocks")
        new_off = np.ctypeslib.as_array((C.c_uint32 * (nb + 1)).from_address(ob.value)).copy()
        n_ids = int(new_off[-1])
        new_ids = np.ctypeslib.as_array((C.c_uint32 * max(1, n_ids)).from_address(oi.value)).copy()[:n_ids]
        self.n_reads = int(nr.value)
        return new_off, new_ids

    def pool_truncate(self, n_reads: int):
        self._check(self._lib.fcx_pool_truncate(self._h, n_reads), "fcx_pool_truncate")
        self.n_reads = n_reads

    def generate_consensus(self, seqs: Sequence[bytes], min_cov: int, min_idt: float, K: int = 8) -> bytes:
        """One seed block given as sequences (seqs[0] = seed), like the reference call."""
        self.upload_pool(seqs)
        return self.consensus_blocks([list(range(len(seqs)))], min_cov, min_idt, K)[0]

    def align_pairs(self, q_ids: Sequence[int], t_ids: Sequence[int], ranges=None, band_tolerance: int = 1500):
        """Batched DWA.align of pool sequences (graph_to_contig.get_aln_data), distance only.
        -> int32 array [n, 4]: aln_str_size, dist, aln_q_e, aln_t_e."""
        q = np.ascontiguousarray(q_ids, dtype=np.uint32)
        t = np.ascontiguousarray(t_ids, dtype=np.uint32)
        r = None if ranges is None else np.ascontiguousarray(ranges, dtype=np.int32).reshape(-1, 4)
        out = np.zeros((q.shape[0], 4), dtype=np.int32)
        self._check(self._lib.fcx_align_pairs(self._h, q.shape[0], q.ctypes.data, t.ctypes.data,
                                              None if r is None else r.ctypes.data, band_tolerance, out.ctypes.data),
                    "fcx_align_pairs")
        return out

    # -- diagnostics --------------------------------------------------------------------
    def pair_info(self) -> List[PairInfo]:
        n = C.c_uint64()
        self._lib.fcx_last_pair_info(self._h, None, 0, C.byref(n))
        arr = (PairInfo * max(1, n.value))()
        self._lib.fcx_last_pair_info(self._h, arr, n.value, C.byref(n))
        return list(arr)[: n.value]

    def stats(self) -> dict:
        t = (C.c_double * len(T_NAMES))()
        c = (C.c_uint64 * len(C_NAMES))()
        self._lib.fcx_last_stats(self._h, t, c)
        d = {"ms_" + k: t[i] for i, k in enumerate(T_NAMES)}
        d.update({k: int(c[i]) for i, k in enumerate(C_NAMES)})
        return d


class MultiEngine(Engine):
    """Several GPUs in one process: one read store on every device, seed blocks sharded in
    contiguous cost-balanced slices, results merged in seed order (fcx_multi_*).
    Same surface as Engine."""

    def __init__(self, devices: Sequence[int]):
        self._lib = lib()
        self._open(devices)

    def _open(self, devices: Sequence[int]):
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        if self._lib.fcx_multi_create(arr, len(devices), C.byref(h)) != 0:
            raise EngineError("fcx_multi_create(%s): %s" % (list(devices), self._lib.fcx_multi_last_error(None).decode()))
        self._h = h
        self.devices = list(devices)
        self.n_reads = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fcx_multi_destroy(self._h)
            self._h = None

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError("%s: %s" % (what, self._lib.fcx_multi_last_error(self._h).decode()))

    def set_option(self, name: str, value: float):
        self._check(self._lib.fcx_multi_set_option(self._h, name.encode(), float(value)), "fcx_multi_set_option")

    def upload_pool_raw(self, bases_ptr: int, offsets: np.ndarray):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.shape[0] - 1
        self._check(self._lib.fcx_multi_pool_upload(self._h, bases_ptr, offsets.ctypes.data, n), "fcx_multi_pool_upload")
        self.n_reads = self._reserved = n

    def peer_bytes(self) -> int:
        return int(self._lib.fcx_multi_peer_bytes(self._h))

    def consensus_blocks_raw(self, block_off: np.ndarray, read_ids: np.ndarray, min_cov: int,
                             min_idt: float, K: int = 8, copy: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        block_off = np.ascontiguousarray(block_off, dtype=np.uint32)
        read_ids = np.ascontiguousarray(read_ids, dtype=np.uint32)
        nb = block_off.shape[0] - 1
        ob, oo = C.c_void_p(), C.c_void_p()
        self._check(self._lib.fcx_multi_consensus_blocks(self._h, nb, block_off.ctypes.data, read_ids.ctypes.data,
                                                         min_cov, K, min_idt, C.byref(ob), C.byref(oo)),
                    "fcx_multi_consensus_blocks")
        off = np.ctypeslib.as_array((C.c_uint64 * (nb + 1)).from_address(oo.value))
        total = int(off[-1])
        data = np.ctypeslib.as_array((C.c_uint8 * total).from_address(ob.value)) if total else np.zeros(0, dtype=np.uint8)
        return (data.copy(), off.copy()) if copy else (data, off)

    def pair_info(self) -> List[PairInfo]:
        n = C.c_uint64()
        self._lib.fcx_multi_last_pair_info(self._h, None, 0, C.byref(n))
        arr = (PairInfo * max(1, n.value))()
        self._lib.fcx_multi_last_pair_info(self._h, arr, n.value, C.byref(n))
        return list(arr)[: n.value]

    def stats(self) -> dict:
        t = (C.c_double * len(T_NAMES))()
        c = (C.c_uint64 * len(C_NAMES))()
        self._lib.fcx_multi_last_stats(self._h, t, c)
        d = {"ms_" + k: t[i] for i, k in enumerate(T_NAMES)}
        d.update({k: int(c[i]) for i, k in enumerate(C_NAMES)})
        return d


class StreamParser:
    """LA4Falcon block-stream parser in the native library (fcx_parser_*), the C counterpart of
    falcon_kit/mains/consensus.py:get_seq_data + get_longest_reads."""

    def __init__(self, min_n_read: int, min_len_aln: int, max_n_read: int, min_cov_aln: int, max_cov_aln: int):
        self._lib = lib()
        self._h = self._lib.fcx_parser_create(min_n_read, min_len_aln, max_n_read, min_cov_aln, max_cov_aln)
        self.stopped = False

    def close(self):
        if self._h:
            self._lib.fcx_parser_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def feed(self, data: bytes, eof: bool = False) -> int:
        rc = self._lib.fcx_parser_feed(self._h, data, len(data), 1 if eof else 0)
        if rc < 0:
            self.stopped = True
            return -rc - 1
        return rc

    def pending(self) -> int:
        return self._lib.fcx_parser_pending(self._h)

    def take(self, max_blocks: int, max_bases: int):
        """-> (bases_ptr, offsets (uint64 array), block_off, read_ids (uint32 arrays), seed ids)"""
        b, o, bo, ri, si = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nr, nb = C.c_uint32(), C.c_uint32()
        self._lib.fcx_parser_take(self._h, max_blocks, max_bases, C.byref(b), C.byref(o), C.byref(nr), C.byref(bo),
                                  C.byref(ri), C.byref(nb), C.byref(si))
        n_reads, n_blocks = nr.value, nb.value
        offsets = np.ctypeslib.as_array((C.c_uint64 * (n_reads + 1)).from_address(o.value)).copy()
        block_off = np.ctypeslib.as_array((C.c_uint32 * (n_blocks + 1)).from_address(bo.value)).copy()
        read_ids = np.ctypeslib.as_array((C.c_uint32 * max(1, int(block_off[-1]))).from_address(ri.value)).copy()[:int(block_off[-1])]
        ids, addr = [], si.value
        for _ in range(n_blocks):
            sid = C.string_at(addr)
            ids.append(sid.decode())
            addr += len(sid) + 1
        return b.value, offsets, block_off, read_ids, ids


class DazzDB:
    """A Dazzler read database + .las files read directly (fcx_dazz_*): replaces
    `LA4Falcon -H$CUTOFF -fo db las | ...` (falcon_kit/mains/consensus_task.py:81-90)."""

    def __init__(self, db_path: str, library: Optional[C.CDLL] = None):
        self._lib = library or lib()
        h = C.c_void_p()
        if self._lib.fcx_dazz_open(os.fsencode(db_path), C.byref(h)) != 0:
            raise EngineError("fcx_dazz_open: %s" % self._lib.fcx_dazz_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fcx_dazz_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s: %s" % (what, self._lib.fcx_dazz_last_error(self._h).decode()))

    @property
    def n_reads(self) -> int:
        return int(self._lib.fcx_dazz_nreads(self._h))

    def upload(self, engine: "Engine"):
        """Every read of the DB into the engine's pool: id 2r forward, 2r + 1 reverse complement."""
        self._check(self._lib.fcx_dazz_upload(self._h, engine._h), "fcx_dazz_upload")
        engine.n_reads = engine._reserved = 2 * self.n_reads

    def open_las(self, las_path: str):
        self._check(self._lib.fcx_las_open(self._h, os.fsencode(las_path)), "fcx_las_open")

    def take(self, seed_cutoff: int, min_n_read: int, min_len_aln: int, max_n_read: int, min_cov_aln: int,
             max_cov_aln: int, max_blocks: int, max_pairs: int):
        """-> (block_off, read_ids, seed_ids, done)"""
        bo, ri, sid, nb, done = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32(), C.c_int()
        self._check(self._lib.fcx_las_take(self._h, seed_cutoff, min_n_read, min_len_aln, max_n_read, min_cov_aln,
                                           max_cov_aln, max_blocks, max_pairs, C.byref(bo), C.byref(ri), C.byref(nb),
                                           C.byref(sid), C.byref(done)), "fcx_las_take")
        n = nb.value
        block_off = np.ctypeslib.as_array((C.c_uint32 * (n + 1)).from_address(bo.value)).copy()
        n_ids = int(block_off[-1])
        read_ids = (np.ctypeslib.as_array((C.c_uint32 * n_ids).from_address(ri.value)).copy() if n_ids
                    else np.zeros(0, dtype=np.uint32))
        ids, p = [], sid.value
        for _ in range(n):
            t = C.string_at(p)
            ids.append(t.decode())
            p += len(t) + 1
        return block_off, read_ids, ids, bool(done.value)
