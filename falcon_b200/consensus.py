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
from typing import List, Sequence

LOG = logging.getLogger()

K = 8  # consensus.py:270


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
    parser.add_argument("--db", type=str, default=None,
                        help="read a Dazzler DB directly (with --las) instead of LA4Falcon text on stdin: the whole read "
                             "store is uploaded once, 2-bit packed, and the .las overlap records are turned into seed "
                             "blocks with LA4Falcon's -f -o -H rules")
    parser.add_argument("--las", action="append", default=None, metavar="FILE", help=".las file(s) for --db, processed in order")
    parser.add_argument("-H", "--seed-cutoff", type=int, default=0, help="LA4Falcon's -H: only reads at least this long are seeds")
    parser.add_argument("--batch-blocks", type=int, default=1024, help="seed blocks per GPU batch")
    parser.add_argument("--batch-bases", type=int, default=1 << 30, help="read bases per GPU batch")
    return parser.parse_args(_normalise_flags(argv[1:]))


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
        if args.trim:                 # get_consensus_with_trim (consensus.py:123-158), on the device
            block_off, read_ids = engine.trim_blocks_raw(block_off, read_ids, args.edge_tolerance, args.trim_size,
                                                         args.max_n_read, args.max_cov_aln)
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


def run_dazz(args, engine, out):
    """--db / --las: Dazzler files in, FASTA out."""
    from .binding import DazzDB
    if not args.las:
        raise SystemExit("--db needs at least one --las")
    if not hasattr(engine, "trim_blocks_raw"):
        raise SystemExit("--db runs on one device: use --device instead of --devices")
    db = DazzDB(args.db, getattr(engine, "_lib", None))
    try:
        db.upload(engine)
        n_store = engine.n_reads
        for las in args.las:
            db.open_las(las)
            done = False
            while not done:
                block_off, read_ids, ids, done = db.take(args.seed_cutoff, args.min_n_read, args.min_len_aln, args.max_n_read,
                                                         args.min_cov_aln, args.max_cov_aln, args.batch_blocks, 1 << 22)
                if not ids:
                    continue
                if args.trim:
                    block_off, read_ids = engine.trim_blocks_raw(block_off, read_ids, args.edge_tolerance, args.trim_size,
                                                                 args.max_n_read, args.max_cov_aln)
                data, off = engine.consensus_blocks_raw(block_off, read_ids, args.min_cov, args.min_idt, K)
                if args.trim:
                    engine.pool_truncate(n_store)
                raw = data.tobytes()
                for i, sid in enumerate(ids):
                    emit(out, raw[int(off[i]):int(off[i + 1])].decode(), sid, args)
    finally:
        db.close()
    out.flush()


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
    if args.trim and not hasattr(engine, "trim_blocks_raw"):
        raise SystemExit("--trim runs on one device: use --device instead of --devices")
    if getattr(args, "db", None):
        return run_dazz(args, engine, stdout)
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


def main(argv=sys.argv):
    args = parse_args(argv)
    run(args)


if __name__ == "__main__":
    main(sys.argv)
