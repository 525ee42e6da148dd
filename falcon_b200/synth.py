"""Seeded synthetic PacBio-CLR-like data for the fc_consensus hot path.

Uniform-random genome, reads sampled uniformly on both strands, per-base error model
ins 9 % / del 4.5 % / sub 1.5 % (15 % total), seed blocks built from ground truth (every read
overlapping the seed by >= ``min_ovl`` bases on the genome, oriented to the seed's strand, whole
read -- what ``LA4Falcon -fo`` hands to ``fc_consensus``; reference consumer:
falcon_kit/mains/consensus.py:161-209).

Two output shapes:
  * a *read pool* (distinct sequences + per-block index lists) for the resident-read-store path;
  * LA4Falcon block text (``"%08d SEQ"`` lines, ``+ +`` after each block, ``- -`` at the end)
    for the drop-in CLI path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_codes(n: int, rng: np.random.Generator) -> np.ndarray:
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def codes_to_bytes(codes: np.ndarray) -> bytes:
    return _ASCII[codes].tobytes()


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    return (3 - codes[::-1]).astype(np.uint8)


def add_errors(tmpl: np.ndarray, rng: np.random.Generator, p_ins: float = 0.09,
               p_del: float = 0.045, p_sub: float = 0.015) -> np.ndarray:
    """Apply the CLR-like error model to a template (uint8 codes)."""
    n = tmpl.shape[0]
    keep = rng.random(n) >= p_del
    out = tmpl.copy()
    sub = rng.random(n) < p_sub
    out[sub] = (out[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    # geometric insertion run after each template base, mean p_ins
    r = p_ins / (1.0 + p_ins)
    n_ins = rng.geometric(1.0 - r, size=n) - 1
    reps = keep.astype(np.int64) + n_ins
    total = int(reps.sum())
    res = np.repeat(out, reps)
    # positions that are inserted bases: everything except the first copy of a kept base
    starts = np.cumsum(reps) - reps
    is_tmpl = np.zeros(total, dtype=bool)
    is_tmpl[starts[keep & (reps > 0)]] = True
    ins_mask = ~is_tmpl
    res[ins_mask] = rng.integers(0, 4, size=int(ins_mask.sum()), dtype=np.uint8)
    return res


@dataclass
class SynthSet:
    """A read pool plus seed blocks that index into it."""
    pool: List[bytes]                      # distinct sequences (ASCII, upper-case ACGT)
    blocks: List[np.ndarray]               # per block: int32 pool indices, [0] = seed
    seed_ids: List[str]                    # printable seed id per block
    streams: List[np.ndarray] = field(default_factory=list)  # per block: LA4Falcon stream order
    read_len: int = 0
    genome_size: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def n_pairs(self) -> int:
        return int(sum(len(b) - 1 for b in self.blocks))

    def block_seqs(self, bi: int) -> List[bytes]:
        return [self.pool[i] for i in self.blocks[bi]]

    def la4falcon_text(self, block_ids: Optional[Sequence[int]] = None) -> bytes:
        """LA4Falcon-style stream for the CLI path.  The first line of each block is the seed
        itself (the parser appends it twice by design, consensus.py:183-190), then the reads."""
        out = []
        ids = range(len(self.blocks)) if block_ids is None else block_ids
        for bi in ids:
            idx = self.streams[bi]
            out.append(b"%s %s\n" % (self.seed_ids[bi].encode(), self.pool[idx[0]]))
            for pi in idx[1:]:
                out.append(b"%08d %s\n" % (90000000 + int(pi), self.pool[pi]))
            out.append(b"+ +\n")
        out.append(b"- -\n")
        return b"".join(out)


@dataclass
class Geometry:
    """Where every read of a synthetic set lies on the genome.  A pure function of the parameters
    and the seed, so every rank of a multi-GPU job can compute it independently."""
    genome: np.ndarray
    starts: np.ndarray
    lens: np.ndarray
    ends: np.ndarray
    strands: np.ndarray
    seed: int
    p_ins: float
    p_del: float
    p_sub: float
    genome_b: Optional[np.ndarray] = None      # second haplotype (diploid sets): read r comes from haplotype hap[r]
    hap: Optional[np.ndarray] = None

    @property
    def n_reads(self) -> int:
        return int(self.starts.shape[0])


def make_geometry(genome_size: int, read_len: int, coverage: float, seed: int = 20260924,
                  len_sigma: float = 0.0, p_ins: float = 0.09, p_del: float = 0.045,
                  p_sub: float = 0.015, ploidy: int = 1, het: float = 0.01) -> Geometry:
    """``ploidy`` 2: two haplotypes differing by ``het`` SNPs (BASELINE config 5); ``coverage`` is the
    total over both, every read is drawn from one haplotype, seed blocks are built from positional
    overlap regardless of haplotype (what an overlapper would report at 1 % divergence)."""
    rng = np.random.default_rng(seed)
    genome = random_codes(genome_size, rng)
    n_reads = max(2, int(round(genome_size * coverage / read_len)))
    if len_sigma > 0:
        mu = np.log(read_len) - 0.5 * len_sigma ** 2
        lens = np.clip(rng.lognormal(mu, len_sigma, n_reads).astype(np.int64), 1000,
                       min(genome_size, 99000))
    else:
        lens = np.full(n_reads, min(read_len, genome_size), dtype=np.int64)
    starts = (rng.random(n_reads) * (genome_size - lens + 1)).astype(np.int64)
    order = np.argsort(starts, kind="stable")
    starts, lens = starts[order], lens[order]
    strands = rng.integers(0, 2, n_reads)
    genome_b = hap = None
    if ploidy == 2:
        rng_b = np.random.default_rng([seed, 0x68617062])
        genome_b = genome.copy()
        for c0 in range(0, genome_size, 1 << 26):                 # in pieces: a 1 Gb genome must not need 8 GB of doubles
            c1 = min(genome_size, c0 + (1 << 26))
            snp = c0 + np.flatnonzero(rng_b.random(c1 - c0) < het)
            genome_b[snp] = (genome_b[snp] + rng_b.integers(1, 4, snp.shape[0]).astype(genome_b.dtype)) & 3
        hap = rng_b.integers(0, 2, n_reads)
    return Geometry(genome, starts, lens, starts + lens, strands, seed, p_ins, p_del, p_sub, genome_b, hap)


def gen_reads(geo: Geometry, r0: int, r1: int) -> List[bytes]:
    """Noisy copies of reads r0..r1-1: two pool entries per read (forward, reverse complement).
    Every read has its own RNG stream keyed by (seed, read index): any rank can generate any slice
    and gets exactly the bytes every other rank would."""
    out: List[bytes] = []
    for r in range(r0, r1):
        rng = np.random.default_rng([geo.seed, r])
        g = geo.genome_b if (geo.hap is not None and geo.hap[r]) else geo.genome
        fwd = add_errors(g[geo.starts[r]:geo.ends[r]], rng, geo.p_ins, geo.p_del, geo.p_sub)
        if fwd.shape[0] > 99998:
            fwd = fwd[:99998]
        out.append(codes_to_bytes(fwd))
        out.append(codes_to_bytes(revcomp_codes(fwd)))
    return out


def build_blocks(geo: Geometry, noisy_len: np.ndarray, seeds: Sequence[int], min_ovl: int = 1000,
                 max_n_read: int = 200) -> Tuple[List[np.ndarray], List[str], List[np.ndarray]]:
    """Seed blocks from ground truth.  Block layout mirrors what get_seq_data + get_longest_reads
    produce (consensus.py:26-45,161-209): ``[seed, seed, reads sorted by -len (stable)]`` capped at
    ``max_n_read`` entries.  ``noisy_len[r]`` = length of read r's noisy copy."""
    starts, ends = geo.starts, geo.ends
    blocks: List[np.ndarray] = []
    seed_ids: List[str] = []
    streams: List[np.ndarray] = []
    max_len = int(geo.lens.max())
    for s in seeds:
        lo = int(np.searchsorted(starts, starts[s] - max_len, side="left"))
        hi = int(np.searchsorted(starts, ends[s], side="left"))
        cand = np.arange(lo, hi)
        ovl = np.minimum(ends[cand], ends[s]) - np.maximum(starts[cand], starts[s])
        members = cand[(ovl >= min_ovl) & (cand != s)]
        # stable sort by -len, as get_longest_reads does on the noisy sequences
        members = members[np.argsort(-noisy_len[members], kind="stable")]
        o = int(geo.strands[s])  # 0: forward copies, 1: reverse-complement copies
        idx = [2 * s + o, 2 * s + o] + [2 * int(m) + o for m in members]
        stream = np.asarray([idx[0]] + idx[2:], dtype=np.int32)
        # the duplicated seed copy takes part in the stable sort too (it is seqs[1])
        rest = idx[1:]
        rest_len = np.array([noisy_len[i >> 1] for i in rest])
        rest = [rest[i] for i in np.argsort(-rest_len, kind="stable")]
        idx = ([idx[0]] + rest)[:max_n_read]
        blocks.append(np.asarray(idx, dtype=np.int32))
        seed_ids.append("%08d" % s)
        streams.append(stream)
    return blocks, seed_ids, streams


def make_set(genome_size: int, read_len: int, coverage: float, seed: int = 20260924,
             n_blocks: Optional[int] = None, min_ovl: int = 1000, max_n_read: int = 200,
             len_sigma: float = 0.0, p_ins: float = 0.09, p_del: float = 0.045,
             p_sub: float = 0.015, block_stride: int = 1) -> SynthSet:
    """Build a synthetic set in one process.  Each read contributes two pool entries
    (forward-strand noisy copy and its reverse complement); a block uses the orientation of its
    seed for every member."""
    geo = make_geometry(genome_size, read_len, coverage, seed, len_sigma, p_ins, p_del, p_sub)
    pool = gen_reads(geo, 0, geo.n_reads)
    noisy_len = np.fromiter((len(pool[2 * r]) for r in range(geo.n_reads)), dtype=np.int64, count=geo.n_reads)
    seeds = list(range(0, geo.n_reads, block_stride))
    if n_blocks is not None:
        seeds = seeds[:n_blocks]
    blocks, seed_ids, streams = build_blocks(geo, noisy_len, seeds, min_ovl, max_n_read)
    return SynthSet(pool=pool, blocks=blocks, seed_ids=seed_ids, streams=streams, read_len=read_len,
                    genome_size=genome_size,
                    meta=dict(seed=seed, coverage=coverage, p_ins=p_ins, p_del=p_del, p_sub=p_sub,
                              min_ovl=min_ovl, max_n_read=max_n_read, n_reads=geo.n_reads))
