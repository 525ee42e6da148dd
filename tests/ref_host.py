"""TEST INFRASTRUCTURE: an independent Python model of the host-side rules of the reference CLI
(falcon_kit/mains/consensus.py) -- the stdin block parser (:161-209), read selection (:26-45) and the
--trim read cutting (:48-99, :123-147).  It is the checker for the native parser (fcx_parser_*), for
fcx_las_take and for the device --trim path (fcx_trim_blocks); it is itself pinned by the golden
output of the unmodified reference CLI (tests/test_oracle_golden.py).  The product (falcon_b200/)
does not import this file.

Public names keep the reference's (the tests read like its own would); the bodies are written from
the rules, which are listed in SURVEY.md 8(a) rows a1 / a2 and section 3.3."""
from typing import Iterator, List, Tuple

MAX_SEQ = 100000          # longer sequences are cut to MAX_SEQ - 1 characters
K = 8


def get_longest_reads(seqs, max_n_read, max_cov_aln, sort=True):
    """seed + the other reads longest first (stable), capped by count and -- if max_cov_aln > 0 -- by
    the coverage they add up to: reads are taken while (bases so far) // seed_len <= max_cov_aln."""
    seed, rest = seqs[:1], list(seqs[1:])
    if sort:
        rest.sort(key=len, reverse=True)              # list.sort is stable; reverse keeps ties in order
    keep = max_n_read
    if max_cov_aln > 0:
        covered, taken = 0, 0
        for s in rest:
            if covered // len(seed[0]) > max_cov_aln:
                break
            taken += 1
            covered += len(s)
        keep = min(1 + taken, max_n_read)
    return (seed + rest)[:keep]


class _Block:
    """The reads of the seed block being collected."""

    def __init__(self):
        self.seqs: List[bytes] = []
        self.seen = set()
        self.seed_id = None
        self.bases = 0

    def add(self, rid: bytes, seq: bytes):
        if not self.seqs:                     # the first read is the seed; it is listed once here ...
            self.seqs.append(seq)
            self.seed_id = rid
        if rid not in self.seen:              # ... and once more as an ordinary read (its id is new)
            self.seen.add(rid)
            self.seqs.append(seq)
            self.bases += len(seq)

    def complete(self, min_n_read, min_cov_aln) -> bool:
        return len(self.seqs) >= min_n_read and self.bases // len(self.seqs[0]) >= min_cov_aln


def get_seq_data(stream, config, min_n_read, min_len_aln) -> Iterator[Tuple[List[bytes], str]]:
    """LA4Falcon text -> (reads of a seed block, seed id).  Only lines of exactly two whitespace-separated
    tokens count; "+" closes a block (kept if it has enough reads and coverage), "*" drops it, "-" ends
    the stream; reads shorter than min_len_aln are skipped."""
    _min_cov, _k, max_n_read, _min_idt, _edge, _trim, min_cov_aln, max_cov_aln = config
    blk = _Block()
    for line in stream:
        tok = line.split()
        if len(tok) != 2:
            continue
        rid, seq = tok
        if len(seq) > MAX_SEQ:
            seq = seq[:MAX_SEQ - 1]
        if rid == b"-":
            return
        if rid == b"*":
            blk = _Block()
        elif rid == b"+":
            if blk.seqs and blk.complete(min_n_read, min_cov_aln):
                yield get_longest_reads(blk.seqs, max_n_read, max_cov_aln, sort=True), blk.seed_id.decode()
            blk = _Block()
        elif len(seq) >= min_len_aln:
            blk.add(rid, seq)


def get_alignment(ref, seq1: bytes, seq0: bytes, edge_tolerance=1000):
    """Where read seq1 maps on seed seq0 according to the masked k-mer chaining (the C side is done by the
    compiled reference, oracle.Ref.trim_range) -> (s1, e1, s0, e0, aln_size, aln_score, "aln" | "none")."""
    none = (0, 0, 0, 0, 0, 0, "none")
    _n, s1, e1, s0, e0, chain = ref.trim_range(seq1, seq0, K)
    e1 = min(e1 + K + K // 2, len(seq1))           # the chain ends on k-mer START positions
    e0 = min(e0 + K + K // 2, len(seq0))
    long_enough = e1 - s1 > 500
    aln_size = max(e1 - s1, e0 - s0) if long_enough else 1
    aln_score = int(chain * 48) if long_enough else 0
    hangs_left = s1 > edge_tolerance and s0 > edge_tolerance
    hangs_right = len(seq1) - e1 > edge_tolerance and len(seq0) - e0 > edge_tolerance
    if hangs_left or hangs_right or not (long_enough and aln_size > 500):
        return none
    return s1, e1, s0, e0, aln_size, aln_score, "aln"


def trim_block(ref, seqs: List[bytes], config) -> List[bytes]:
    """--trim: every read is cut to its mapped range minus trim_size on both sides; reads that do not map
    well are dropped; the rest follow the seed, longest cut first."""
    _min_cov, _k, max_n_read, _min_idt, edge_tolerance, trim_size, _min_cov_aln, max_cov_aln = config
    seed = seqs[0]
    cuts = []
    for seq in seqs[1:]:
        s1, e1, _s0, _e0, _size, score, status = get_alignment(ref, seq, seed, edge_tolerance)
        if status == "aln" and score > 1000 and e1 - s1 > 500:
            a, b = s1 + trim_size, e1 - trim_size
            cuts.append((b - a, seq[a:b]))
    cuts.sort(key=lambda c: c[0], reverse=True)
    out = [seed] + [c[1] for c in cuts]
    if len(out) - 1 > max_n_read:
        out = get_longest_reads(out, max_n_read, max_cov_aln, sort=False)
    return out
