"""TEST INFRASTRUCTURE: a Python restatement of the reference CLI's host-side logic
(falcon_kit/mains/consensus.py) -- stdin parser, read selection, --trim read cutting -- used as the
checker for the native parser (fcx_parser_*) and for the device --trim path (fcx_trim_blocks).
The product (falcon_b200/) does not import this file."""
from typing import Iterator, List, Tuple


def get_longest_reads(seqs, max_n_read, max_cov_aln, sort=True):
    """Keep the seed plus the longest reads -- consensus.py:26-45 (stable sort by -len)."""
    if sort:
        seqs = seqs[:1] + sorted(seqs[1:], key=lambda x: -len(x))
    longest_n_reads = max_n_read
    if max_cov_aln > 0:
        longest_n_reads = 1
        seed_len = len(seqs[0])
        read_cov = 0
        for seq in seqs[1:]:
            if read_cov // seed_len > max_cov_aln:
                break
            longest_n_reads += 1
            read_cov += len(seq)
        longest_n_reads = min(longest_n_reads, max_n_read)
    return seqs[:longest_n_reads]


def get_seq_data(stream, config, min_n_read, min_len_aln) -> Iterator[Tuple[List[bytes], str]]:
    """Parse the LA4Falcon stream into seed blocks -- consensus.py:161-209.

    ``stream`` yields bytes lines.  Lines that do not split into exactly two tokens are ignored;
    sequences longer than 100000 are cut to 99999; the first read is the seed and is appended
    twice by design; duplicate ids are dropped; ``+`` emits, ``*`` discards, ``-`` stops.
    """
    max_len = 100000
    min_cov, _K, max_n_read, min_idt, edge_tolerance, trim_size, min_cov_aln, max_cov_aln = config
    seqs: List[bytes] = []
    seed_id = None
    seed_len = 0
    read_cov = 0
    read_ids = set()
    for raw in stream:
        l = raw.strip().split()
        if len(l) != 2:
            continue
        read_id, seq = l[0], l[1]
        if len(seq) > max_len:
            seq = seq[:max_len - 1]
        if read_id not in (b"+", b"-", b"*"):
            if len(seq) >= min_len_aln:
                if len(seqs) == 0:
                    seqs.append(seq)  # the "seed"
                    seed_len = len(seq)
                    seed_id = read_id
                if read_id not in read_ids:  # the seed is used again here by design
                    seqs.append(seq)
                    read_ids.add(read_id)
                    read_cov += len(seq)
        elif read_id == b"+":
            if len(seqs) >= min_n_read and read_cov // seed_len >= min_cov_aln:
                seqs = get_longest_reads(seqs, max_n_read, max_cov_aln, sort=True)
                yield seqs, seed_id.decode()
            seqs = []
            read_ids = set()
            seed_id = None
            read_cov = 0
        elif read_id == b"*":
            seqs = []
            read_ids = set()
            seed_id = None
            read_cov = 0
        elif read_id == b"-":
            break


def get_alignment(ref, seq1: bytes, seq0: bytes, edge_tolerance=1000):
    """consensus.py:48-99 with the C calls done by the compiled reference (oracle.Ref.trim_range)."""
    K = 8
    _n, s1, e1, s0, e0, km_score = ref.trim_range(seq1, seq0, K)
    e1 += K + K // 2
    e0 += K + K // 2
    len_1, len_0 = len(seq1), len(seq0)
    e1 = min(e1, len_1)
    e0 = min(e0, len_0)
    aln_size = 1
    aln_score = 0
    if e1 - s1 > 500:
        aln_size = max(e1 - s1, e0 - s0)
        aln_score = int(km_score * 48)
    if s1 > edge_tolerance and s0 > edge_tolerance:
        return 0, 0, 0, 0, 0, 0, "none"
    if len_1 - e1 > edge_tolerance and len_0 - e0 > edge_tolerance:
        return 0, 0, 0, 0, 0, 0, "none"
    if e1 - s1 > 500 and aln_size > 500:
        return s1, e1, s0, e0, aln_size, aln_score, "aln"
    return 0, 0, 0, 0, 0, 0, "none"


def trim_block(ref, seqs: List[bytes], config) -> List[bytes]:
    """consensus.py:123-147."""
    min_cov, _K, max_n_read, min_idt, edge_tolerance, trim_size, min_cov_aln, max_cov_aln = config
    trim_seqs = []
    seed = seqs[0]
    for seq in seqs[1:]:
        s1, e1, s2, e2, aln_size, aln_score, c_status = get_alignment(ref, seq, seed, edge_tolerance)
        if c_status == "none":
            continue
        if aln_score > 1000 and e1 - s1 > 500:
            e1 -= trim_size
            s1 += trim_size
            trim_seqs.append((e1 - s1, seq[s1:e1]))
    trim_seqs.sort(key=lambda x: -x[0])
    out = [seed] + [x[1] for x in trim_seqs]
    if len(out[1:]) > max_n_read:
        out = get_longest_reads(out, max_n_read, max_cov_aln, sort=False)
    return out
