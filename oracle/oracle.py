"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY (see fc_oracle.h).

* ``Ref``    -- oracle/_ref/falcon.so: the UNMODIFIED reference sources (src/c/*.c) compiled by
               oracle/Makefile; struct mirrors follow falcon_kit/falcon_kit.py:19-41,86-106.
* ``Oracle`` -- oracle/_build/liboracle.so: our C restatement (fc_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "falcon.so")
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")


def build(quiet: bool = True) -> None:
    """Compile the restatement (and the reference .so when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class PairInfo(C.Structure):
    _fields_ = [("n_match", C.c_int), ("s1", C.c_int), ("e1", C.c_int), ("s2", C.c_int),
                ("e2", C.c_int), ("score", C.c_long), ("passed_filter", C.c_int),
                ("aligned", C.c_int), ("dist", C.c_int), ("aln_size", C.c_int),
                ("q_e", C.c_int), ("t_e", C.c_int), ("accepted", C.c_int), ("n_tags", C.c_int),
                ("trace_cells", C.c_long)]

    def as_tuple(self):
        return tuple(getattr(self, f) for f, _ in self._fields_)


class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build()
        self.lib = lib = C.CDLL(path)
        lib.orc_generate_consensus.argtypes = [C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_uint,
                                               C.c_double, C.POINTER(PairInfo),
                                               C.POINTER(C.POINTER(C.c_int))]
        lib.orc_generate_consensus.restype = C.c_void_p
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_kmer_range.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_uint,
                                       C.POINTER(PairInfo)]
        lib.orc_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.c_char_p, C.c_char_p, C.POINTER(C.c_long)]
        lib.orc_align.restype = C.c_int

    def generate_consensus(self, seqs: Sequence[bytes], min_cov: int, min_idt: float, K: int = 8,
                           want_info: bool = False, want_eqv: bool = False):
        n = len(seqs)
        arr = (C.c_char_p * n)(*seqs)
        info = (PairInfo * n)() if want_info else None
        eqv_p = C.POINTER(C.c_int)()
        p = self.lib.orc_generate_consensus(arr, n, min_cov, K, min_idt, info,
                                            C.byref(eqv_p) if want_eqv else None)
        cns = C.string_at(p)
        self.lib.orc_free(p)
        out = [cns]
        if want_info:
            out.append(list(info))
        if want_eqv:
            out.append([eqv_p[i] for i in range(len(cns))])
            self.lib.orc_free(C.cast(eqv_p, C.c_void_p))
        return out[0] if len(out) == 1 else tuple(out)

    def kmer_range(self, read: bytes, seed: bytes, K: int = 8) -> PairInfo:
        pi = PairInfo()
        self.lib.orc_kmer_range(read, len(read), seed, len(seed), K, C.byref(pi))
        return pi

    def align(self, q: bytes, t: bytes, band_tolerance: int = 150):
        qa = C.create_string_buffer(len(q) + len(t) + 1)
        ta = C.create_string_buffer(len(q) + len(t) + 1)
        dist, qe, te, cells = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        n = self.lib.orc_align(q, len(q), t, len(t), band_tolerance, C.byref(dist), C.byref(qe),
                               C.byref(te), qa, ta, C.byref(cells))
        return dict(aln_str_size=n, dist=dist.value, q_e=qe.value, t_e=te.value,
                    q_aln=qa.raw[:n], t_aln=ta.raw[:n], cells=cells.value)


# ---- mirrors of the reference structs (falcon_kit/falcon_kit.py:19-41,86-106) -------------
class _KmerLookup(C.Structure):
    _fields_ = [("start", C.c_int), ("last", C.c_int), ("count", C.c_int)]


class _KmerMatch(C.Structure):
    _fields_ = [("count", C.c_int), ("query_pos", C.POINTER(C.c_int)),
                ("target_pos", C.POINTER(C.c_int))]


class _AlnRange(C.Structure):
    _fields_ = [("s1", C.c_int), ("e1", C.c_int), ("s2", C.c_int), ("e2", C.c_int),
                ("score", C.c_long)]


class _ConsensusData(C.Structure):
    _fields_ = [("sequence", C.c_void_p), ("eqv", C.POINTER(C.c_int))]


class _Alignment(C.Structure):
    _fields_ = [("aln_str_size", C.c_int), ("dist", C.c_int), ("aln_q_s", C.c_int),
                ("aln_q_e", C.c_int), ("aln_t_s", C.c_int), ("aln_t_e", C.c_int),
                ("q_aln_str", C.c_void_p), ("t_aln_str", C.c_void_p)]


class Ref:
    """The compiled, unmodified reference (oracle/_ref/falcon.so)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (reference tree absent and no prebuilt copy)")
        self.lib = lib = C.CDLL(path)
        lib.generate_consensus.argtypes = [C.POINTER(C.c_char_p), C.c_uint, C.c_uint, C.c_uint,
                                           C.c_double]
        lib.generate_consensus.restype = C.POINTER(_ConsensusData)
        lib.free_consensus_data.argtypes = [C.POINTER(_ConsensusData)]
        lib.allocate_kmer_lookup.argtypes = [C.c_int]
        lib.allocate_kmer_lookup.restype = C.POINTER(_KmerLookup)
        lib.allocate_seq.argtypes = [C.c_int]
        lib.allocate_seq.restype = C.POINTER(C.c_uint8)
        lib.allocate_seq_addr.argtypes = [C.c_int]
        lib.allocate_seq_addr.restype = C.POINTER(C.c_int)
        lib.add_sequence.argtypes = [C.c_int, C.c_uint, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_uint8), C.POINTER(_KmerLookup)]
        lib.find_kmer_pos_for_seq.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.POINTER(C.c_int),
                                              C.POINTER(_KmerLookup)]
        lib.find_kmer_pos_for_seq.restype = C.POINTER(_KmerMatch)
        lib.find_best_aln_range.argtypes = [C.POINTER(_KmerMatch), C.c_int, C.c_int, C.c_int]
        lib.find_best_aln_range.restype = C.POINTER(_AlnRange)
        lib.find_best_aln_range2.argtypes = [C.POINTER(_KmerMatch), C.c_int, C.c_int, C.c_int]
        lib.find_best_aln_range2.restype = C.POINTER(_AlnRange)
        lib.mask_k_mer.argtypes = [C.c_int, C.POINTER(_KmerLookup), C.c_int]
        lib.free_kmer_match.argtypes = [C.POINTER(_KmerMatch)]
        lib.free_aln_range.argtypes = [C.POINTER(_AlnRange)]
        lib.free_kmer_lookup.argtypes = [C.POINTER(_KmerLookup)]
        lib.free_seq_array.argtypes = [C.POINTER(C.c_uint8)]
        lib.free_seq_addr_array.argtypes = [C.POINTER(C.c_int)]
        lib.align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.align.restype = C.POINTER(_Alignment)
        lib.free_alignment.argtypes = [C.POINTER(_Alignment)]

    def generate_consensus(self, seqs: Sequence[bytes], min_cov: int, min_idt: float, K: int = 8,
                           want_eqv: bool = False):
        n = len(seqs)
        arr = (C.c_char_p * n)(*seqs)
        p = self.lib.generate_consensus(arr, n, min_cov, K, min_idt)
        cns = C.string_at(p[0].sequence)
        eqv = [p[0].eqv[i] for i in range(len(cns))] if want_eqv else None
        self.lib.free_consensus_data(p)
        return (cns, eqv) if want_eqv else cns

    def kmer_range(self, read: bytes, seed: bytes, K: int = 8):
        lib = self.lib
        lk = lib.allocate_kmer_lookup(1 << (2 * K))
        sa = lib.allocate_seq(len(seed))
        sda = lib.allocate_seq_addr(len(seed))
        lib.add_sequence(0, K, seed, len(seed), sda, sa, lk)
        km = lib.find_kmer_pos_for_seq(read, len(read), K, sda, lk)
        n_match = km[0].count
        ar = lib.find_best_aln_range(km, K, K * 6, 5)
        res = (n_match, ar[0].s1, ar[0].e1, ar[0].s2, ar[0].e2, ar[0].score)
        lib.free_aln_range(ar)
        lib.free_kmer_match(km)
        lib.free_seq_addr_array(sda)
        lib.free_seq_array(sa)
        lib.free_kmer_lookup(lk)
        return res

    def trim_range(self, read: bytes, seed: bytes, K: int = 8):
        """The C calls of get_alignment (falcon_kit/mains/consensus.py:50-63): masked k-mer hits and
        find_best_aln_range2(K, K * 50, 25) -> (n_match, s1, e1, s2, e2, score), raw."""
        lib = self.lib
        lk = lib.allocate_kmer_lookup(1 << (2 * K))
        sa = lib.allocate_seq(len(seed))
        sda = lib.allocate_seq_addr(len(seed))
        lib.add_sequence(0, K, seed, len(seed), sda, sa, lk)
        lib.mask_k_mer(1 << (2 * K), lk, 16)
        km = lib.find_kmer_pos_for_seq(read, len(read), K, sda, lk)
        n_match = km[0].count
        ar = lib.find_best_aln_range2(km, K, K * 50, 25)
        res = (n_match, ar[0].s1, ar[0].e1, ar[0].s2, ar[0].e2, ar[0].score)
        lib.free_aln_range(ar)
        lib.free_kmer_match(km)
        lib.free_seq_addr_array(sda)
        lib.free_seq_array(sa)
        lib.free_kmer_lookup(lk)
        return res

    def align(self, q: bytes, t: bytes, band_tolerance: int = 150):
        a = self.lib.align(q, len(q), t, len(t), band_tolerance, 1)
        n = a[0].aln_str_size
        res = dict(aln_str_size=n, dist=a[0].dist, q_e=a[0].aln_q_e, t_e=a[0].aln_t_e,
                   q_aln=C.string_at(a[0].q_aln_str, n), t_aln=C.string_at(a[0].t_aln_str, n))
        self.lib.free_alignment(a)
        return res
