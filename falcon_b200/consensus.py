"""fc_consensus on B200 -- drop-in for ``python -m falcon_kit.mains.consensus``.

Same stdin (LA4Falcon block text), same options, same FASTA on stdout as the reference CLI
(falcon_kit/mains/consensus.py).  What changes is the engine: instead of a
``multiprocessing.Pool`` of workers calling ``falcon.generate_consensus`` once per seed block
(consensus.py:264-274), parsed blocks are batched and handed to the GPU engine
(``fcx_consensus_blocks``), and results are printed in stdin order (the ordering contract of
``imap``).

    LA4Falcon -H$CUTOFF -fo db las | python -m falcon_b200.consensus --output-multi --min-idt 0.70 \
        --min-cov 4 --max-n-read 200 --n-core 24 > cns.fasta
"""
from __future__ import annotations

import argparse
import logging
import re
import sys
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

LOG = logging.getLogger()

K = 8  # consensus.py:270


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


def get_alignment(seq1: bytes, seq0: bytes, edge_tolerance=1000):
    """K-mer range of seq1 on seq0 for the --trim path -- consensus.py:48-99."""
    from . import binding
    kup = binding.kup()
    lk_ptr = kup.allocate_kmer_lookup(1 << (K * 2))
    sa_ptr = kup.allocate_seq(len(seq0))
    sda_ptr = kup.allocate_seq_addr(len(seq0))
    kup.add_sequence(0, K, seq0, len(seq0), sda_ptr, sa_ptr, lk_ptr)
    kup.mask_k_mer(1 << (K * 2), lk_ptr, 16)
    kmer_match_ptr = kup.find_kmer_pos_for_seq(seq1, len(seq1), K, sda_ptr, lk_ptr)
    aln_range_ptr = kup.find_best_aln_range2(kmer_match_ptr, K, K * 50, 25)
    aln_range = aln_range_ptr[0]
    kup.free_kmer_match(kmer_match_ptr)
    s1, e1, s0, e0, km_score = aln_range.s1, aln_range.e1, aln_range.s2, aln_range.e2, aln_range.score
    e1 += K + K // 2
    e0 += K + K // 2
    kup.free_aln_range(aln_range_ptr)
    len_1, len_0 = len(seq1), len(seq0)
    e1 = min(e1, len_1)
    e0 = min(e0, len_0)
    aln_size = 1
    aln_score = 0
    if e1 - s1 > 500:
        aln_size = max(e1 - s1, e0 - s0)
        aln_score = int(km_score * 48)
    kup.free_seq_addr_array(sda_ptr)
    kup.free_seq_array(sa_ptr)
    kup.free_kmer_lookup(lk_ptr)
    if s1 > edge_tolerance and s0 > edge_tolerance:
        return 0, 0, 0, 0, 0, 0, "none"
    if len_1 - e1 > edge_tolerance and len_0 - e0 > edge_tolerance:
        return 0, 0, 0, 0, 0, 0, "none"
    if e1 - s1 > 500 and aln_size > 500:
        return s1, e1, s0, e0, aln_size, aln_score, "aln"
    return 0, 0, 0, 0, 0, 0, "none"


def trim_block(seqs: List[bytes], config) -> List[bytes]:
    """Read trimming of get_consensus_with_trim -- consensus.py:123-147."""
    min_cov, _K, max_n_read, min_idt, edge_tolerance, trim_size, min_cov_aln, max_cov_aln = config
    trim_seqs = []
    seed = seqs[0]
    for seq in seqs[1:]:
        s1, e1, s2, e2, aln_size, aln_score, c_status = get_alignment(seq, seed, edge_tolerance)
        if c_status == "none":
            continue
        if aln_score > 1000 and e1 - s1 > 500:
            e1 -= trim_size
            s1 += trim_size
            trim_seqs.append((e1 - s1, seq[s1:e1]))
    trim_seqs.sort(key=lambda x: -x[0])  # use longest alignment first
    out = [seed] + [x[1] for x in trim_seqs]
    if len(out[1:]) > max_n_read:
        out = get_longest_reads(out, max_n_read, max_cov_aln, sort=False)
    return out


def format_seq(seq, col):
    return "\n".join([seq[i:(i + col)] for i in range(0, len(seq), col)])


def _normalise_flags(argv: Sequence[str]) -> List[str]:
    """Accept the underscore spellings the pipeline used to pass (falcon_kit/functional.py:403-417)."""
    out = []
    for a in argv:
        if a.startswith("--") and "_" in a.split("=", 1)[0]:
            head, sep, tail = a.partition("=")
            a = head.replace("_", "-") + sep + tail
        out.append(a)
    return out


def parse_devices(spec: str):
    """'0-3' / '0,2,5' / '1' -> list of device ordinals."""
    out = []
    for part in spec.split(","):
        part = part.strip()
        if "-" in part:
            a, b = part.split("-", 1)
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    if not out:
        raise ValueError("--devices: empty device list")
    return out


def parse_args(argv):
    """Same options and defaults as consensus.py:216-251."""
    parser = argparse.ArgumentParser(
        description="a B200-native consensus sequence generator (drop-in for fc_consensus)",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--n-core", type=int, default=24,
                        help="accepted for compatibility; the GPU engine replaces the process pool")
    parser.add_argument("--min-cov", type=int, default=6, help="minimum coverage to break the consensus")
    parser.add_argument("--min-cov-aln", type=int, default=10,
                        help="minimum coverage of alignment data; a seed read with less than MIN_COV_ALN "
                             "average depth of coverage will be completely ignored")
    parser.add_argument("--max-cov-aln", type=int, default=0,
                        help="maximum coverage of alignment data; excess shorter alignments will be ignored")
    parser.add_argument("--min-len-aln", type=int, default=0,
                        help="minimum length of a sequence in an alignment to be used in consensus")
    parser.add_argument("--min-n-read", type=int, default=10,
                        help="1 + minimum number of reads used in generating the consensus")
    parser.add_argument("--max-n-read", type=int, default=500,
                        help="1 + maximum number of reads used in generating the consensus")
    parser.add_argument("--trim", action="store_true", default=False,
                        help="trim the input sequence with k-mer spare dynamic programming to find the mapped range")
    parser.add_argument("--output-full", action="store_true", default=False, help="output uncorrected regions too")
    parser.add_argument("--output-multi", action="store_true", default=False, help="output multi correct regions")
    parser.add_argument("--min-idt", type=float, default=0.70,
                        help="minimum identity of the alignments used for correction")
    parser.add_argument("--edge-tolerance", type=int, default=1000,
                        help="for trimming, the there is unaligned edge leng > edge_tolerance, ignore the read")
    parser.add_argument("--trim-size", type=int, default=50,
                        help="the size for triming both ends from initial sparse aligned region")
    parser.add_argument("-v", "--verbose-level", type=float, default=2.0,
                        help="logging level (WARNING=3, INFO=2, DEBUG=1)")
    # B200 engine knobs (not in the reference)
    parser.add_argument("--device", type=int, default=None, help="CUDA device ordinal (default: $FCX_DEVICE or 0)")
    parser.add_argument("--devices", type=str, default=None,
                        help="several GPUs in this one process, e.g. 0-7 or 0,2,3: one read store on every device, "
                             "seed blocks sharded across them, output in input order")
    parser.add_argument("--stream", action="append", default=None, metavar="IN:OUT",
                        help="read LA4Falcon text from IN and write its FASTA to OUT instead of stdin/stdout; repeat the "
                             "option to serve several producers (files or named pipes) with this one GPU process")
    parser.add_argument("--batch-blocks", type=int, default=1024, help="seed blocks per GPU batch")
    parser.add_argument("--batch-bases", type=int, default=1 << 30, help="read bases per GPU batch")
    parser.add_argument("--python-parser", action="store_true", default=False,
                        help="parse stdin with the Python restatement of get_seq_data instead of the native parser")
    return parser.parse_args(_normalise_flags(argv[1:]))


class BatchRunner:
    """Accumulates parsed seed blocks and runs them through the engine in batches."""

    def __init__(self, engine, min_cov: int, min_idt: float, batch_blocks: int, batch_bases: int):
        self.engine = engine
        self.min_cov, self.min_idt = min_cov, min_idt
        self.batch_blocks, self.batch_bases = batch_blocks, batch_bases
        self.pool: List[bytes] = []
        self.blocks: List[List[int]] = []
        self.ids: List[str] = []
        self.bases = 0

    def add(self, seqs: List[bytes], seed_id: str):
        base = len(self.pool)
        self.pool.extend(seqs)
        self.blocks.append(list(range(base, base + len(seqs))))
        self.ids.append(seed_id)
        self.bases += sum(len(s) for s in seqs)

    def full(self) -> bool:
        return len(self.blocks) >= self.batch_blocks or self.bases >= self.batch_bases

    def flush(self) -> List[Tuple[str, str]]:
        if not self.blocks:
            return []
        self.engine.upload_pool(self.pool)
        cns = self.engine.consensus_blocks(self.blocks, self.min_cov, self.min_idt, K)
        res = [(c.decode(), sid) for c, sid in zip(cns, self.ids)]
        self.pool, self.blocks, self.ids, self.bases = [], [], [], 0
        return res


def emit(out, cns: str, seed_id: str, args, good_region=re.compile("[ACGT]+")):
    """Output stage -- consensus.py:275-299."""
    if len(cns) < 500:
        return
    if args.output_full:
        out.write(">" + seed_id + "_f\n")
        out.write(cns + "\n")
        return
    runs = good_region.findall(cns)
    if len(runs) == 0:
        return
    if args.output_multi:
        seq_i = 0
        for cns_seq in runs:
            if len(cns_seq) < 500:
                continue
            if seq_i >= 10:
                break
            out.write(">prolog/%s%01d/%d_%d\n" % (seed_id, seq_i, 0, len(cns_seq)))
            out.write(format_seq(cns_seq, 80) + "\n")
            seq_i += 1
    else:
        runs.sort(key=lambda x: len(x))
        out.write(">" + seed_id + "\n")
        out.write(runs[-1] + "\n")


def _stream_reader(args, stream_no, fin, q):
    """Producer thread of one input stream: stdin bytes -> native parser (fcx_parser_*, the GIL is
    released inside the C calls) -> batches on the queue.  The parser hands out two alternating
    buffer sets, and the queue holds one batch per stream, so parsing batch n+1 overlaps the GPU
    work on batch n."""
    from .binding import StreamParser
    ps = StreamParser(args.min_n_read, args.min_len_aln, args.max_n_read, args.min_cov_aln, args.max_cov_aln)
    try:
        done = False
        while not done:
            chunk = fin.read(1 << 24)
            eof = not chunk
            pending = ps.feed(chunk or b"", eof)
            done = eof or ps.stopped
            while pending >= args.batch_blocks or (done and pending > 0):
                batch = ps.take(args.batch_blocks, args.batch_bases)
                ev = threading.Event()
                q.put((stream_no, batch, ev))
                ev.wait()             # consumed: its buffer set may be overwritten by the take after next
                pending = ps.pending()
        q.put((stream_no, None, None))
    except BaseException as e:  # noqa: BLE001 -- reported by the consumer
        q.put((stream_no, e, None))
    # the parser (and the buffers of the last batch) must outlive the consumer: closed by the caller
    return ps


def run_native_parser(args, streams, engine):
    """LA4Falcon text -> native parser -> engine, in batches; output in input order per stream.
    `streams` is a list of (binary input file, text output file).  One producer thread per stream
    parses while this thread drives the GPU(s) and writes the FASTA."""
    import queue
    q = queue.Queue(maxsize=max(2, len(streams)))
    parsers = [None] * len(streams)

    def producer(i):
        parsers[i] = _stream_reader(args, i, streams[i][0], q)

    threads = [threading.Thread(target=producer, args=(i,), daemon=True) for i in range(len(streams))]
    for t in threads:
        t.start()
    live = len(streams)
    while live:
        stream_no, batch, ev = q.get()
        if batch is None:
            live -= 1
            continue
        if isinstance(batch, BaseException):
            raise batch
        bases_ptr, offsets, block_off, read_ids, ids = batch
        engine.upload_pool_raw(bases_ptr, offsets)
        ev.set()                      # the read bytes are on the device: the producer may go on
        data, off = engine.consensus_blocks_raw(block_off, read_ids, args.min_cov, args.min_idt, K)
        raw = data.tobytes()
        out = streams[stream_no][1]
        for i, sid in enumerate(ids):
            emit(out, raw[int(off[i]):int(off[i + 1])].decode(), sid, args)
    for t in threads:
        t.join()
    for ps in parsers:
        if ps is not None:
            ps.close()


def run(args, stdin=None, stdout=None, engine=None):
    logging.basicConfig(level=int(round(10 * args.verbose_level)))
    stdin = stdin if stdin is not None else sys.stdin.buffer
    stdout = stdout if stdout is not None else sys.stdout
    if engine is None:
        import os
        from .binding import Engine
        devs = parse_devices(args.devices) if getattr(args, "devices", None) else None
        if devs and len(devs) > 1:
            from .binding import MultiEngine
            engine = MultiEngine(devs)
        else:
            dev = devs[0] if devs else (args.device if args.device is not None else int(os.environ.get("FCX_DEVICE", "0")))
            engine = Engine(dev)
    if not args.trim and not args.python_parser and hasattr(engine, "upload_pool_raw"):
        streams, opened = [(stdin, stdout)], []
        if getattr(args, "stream", None):
            # several LA4Falcon producers feeding this one process (which owns the GPUs):
            # --stream IN:OUT, repeated; IN / OUT are files or named pipes
            streams = []
            for spec in args.stream:
                fin_name, fout_name = spec.split(":", 1)
                fin, fout = open(fin_name, "rb"), open(fout_name, "w")
                opened += [fin, fout]
                streams.append((fin, fout))
        try:
            run_native_parser(args, streams, engine)
        finally:
            for _, fout in streams:
                fout.flush()
            for f in opened:
                f.close()
        return
    config = (args.min_cov, K, args.max_n_read, args.min_idt, args.edge_tolerance, args.trim_size,
              args.min_cov_aln, args.max_cov_aln)
    runner = BatchRunner(engine, args.min_cov, args.min_idt, args.batch_blocks, args.batch_bases)
    for seqs, seed_id in get_seq_data(stdin, config, args.min_n_read, args.min_len_aln):
        if args.trim:
            seqs = trim_block(seqs, config)
        elif len(seqs) > args.max_n_read:          # consensus.py:107-108
            seqs = get_longest_reads(seqs, args.max_n_read, args.max_cov_aln, sort=True)
        runner.add(seqs, seed_id)
        if runner.full():
            for cns, sid in runner.flush():
                emit(stdout, cns, sid, args)
    for cns, sid in runner.flush():
        emit(stdout, cns, sid, args)
    stdout.flush()


def main(argv=sys.argv):
    args = parse_args(argv)
    run(args)


if __name__ == "__main__":
    main(sys.argv)
