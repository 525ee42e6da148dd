#!/usr/bin/env python
"""Throughput of the DROP-IN command: LA4Falcon block text -> `python -m falcon_b200.consensus` -> FASTA
(falcon_kit/mains/consensus_task.py:90: `LA4Falcon ... | python -m falcon_kit.mains.consensus ... > out`).

    python tools/bench_text.py [--blocks 2048] [--streams 1,4,16] [--devices 0]

Writes the synthetic E. coli-like workload as LA4Falcon text into S files under /dev/shm (each pair
re-ships a whole 15 kb read, as LA4Falcon -fo does), runs the CLI in-process with S --stream IN:OUT
pairs (S parser threads feeding the one GPU loop) and reports pairs/s and text GB/s per S.
PARITY GATE: the FASTA of a sample of blocks is compared byte for byte with the same host logic
driven by the reference's own C code (oracle/_ref/falcon.so) -- this is test/bench infrastructure,
the product path never touches oracle/.
One JSON line per S is printed.
"""
from __future__ import annotations

import argparse
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--read-len", type=int, default=15000)
    ap.add_argument("--cov", type=float, default=50)
    ap.add_argument("--blocks", type=int, default=2048)
    ap.add_argument("--streams", default="1,4,16")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--parity-blocks", type=int, default=8)
    ap.add_argument("--tmp", default="/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    a = ap.parse_args()
    from falcon_b200 import consensus, synth

    n_reads = int(round(a.genome * a.cov / a.read_len))
    S = synth.make_set(a.genome, a.read_len, a.cov, n_blocks=a.blocks, max_n_read=200,
                       block_stride=max(1, n_reads // a.blocks))
    n_pairs = S.n_pairs
    opts = ["--output-multi", "--min-idt", "0.70", "--min-cov", "4", "--max-n-read", "200"]
    # ---- parity reference for a few blocks: same host logic, arithmetic by the reference C code
    from helpers import OracleEngine
    from oracle.oracle import Ref, REF_SO, Oracle
    ref_engine = OracleEngine(Ref() if os.path.exists(REF_SO) else Oracle())
    pb = list(range(min(a.parity_blocks, len(S.blocks))))
    want = io.StringIO()
    consensus.run(consensus.parse_args(["consensus"] + opts), stdin=io.BytesIO(S.la4falcon_text(pb)), stdout=want,
                  engine=ref_engine)
    from falcon_b200.binding import Engine, MultiEngine
    devs = consensus.parse_devices(a.devices)
    engine = MultiEngine(devs) if len(devs) > 1 else Engine(devs[0])
    got = io.StringIO()
    consensus.run(consensus.parse_args(["consensus"] + opts), stdin=io.BytesIO(S.la4falcon_text(pb)), stdout=got,
                  engine=engine)
    if got.getvalue() != want.getvalue():
        print(json.dumps({"error": "PARITY GATE FAILED: CLI output differs from the reference C code on %d blocks" % len(pb)}))
        return 3
    # ---- timed runs
    for s_txt in a.streams.split(","):
        ns = int(s_txt)
        per = (len(S.blocks) + ns - 1) // ns
        files, text_bytes = [], 0
        for i in range(ns):
            ids = list(range(i * per, min(len(S.blocks), (i + 1) * per)))
            if not ids:
                continue
            fin = os.path.join(a.tmp, "fcx_text_%d_%d.in" % (os.getpid(), i))
            fout = os.path.join(a.tmp, "fcx_text_%d_%d.out" % (os.getpid(), i))
            t = S.la4falcon_text(ids)
            text_bytes += len(t)
            with open(fin, "wb") as f:
                f.write(t)
            files.append((fin, fout))
        argv = ["consensus"] + opts
        for fin, fout in files:
            argv += ["--stream", "%s:%s" % (fin, fout)]
        args = consensus.parse_args(argv)
        consensus.run(args, engine=engine)                       # warm-up (page cache, device buffers)
        t0 = time.perf_counter()
        consensus.run(args, engine=engine)
        dt = time.perf_counter() - t0
        out_bytes = sum(os.path.getsize(fo) for _, fo in files)
        for fin, fout in files:
            os.unlink(fin); os.unlink(fout)
        print(json.dumps({"metric": "aligned read-pairs/sec fc_consensus (drop-in text path)", "value": n_pairs / dt,
                          "unit": "pairs/s", "streams": len(files), "devices": devs, "blocks": len(S.blocks),
                          "pairs": n_pairs, "text_gb": text_bytes / 1e9, "text_gb_per_s": text_bytes / 1e9 / dt,
                          "fasta_mb": out_bytes / 1e6, "seconds": dt,
                          "parity_gate": "CLI == reference C code on %d blocks" % len(pb)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
