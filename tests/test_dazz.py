"""Dazzler DB + .las input (fcx_dazz_*, --db/--las) against the LA4Falcon text path on the same overlaps.
The files are written by tests/dazz_writer.py (the real tools are absent); what is
checked is that the binary path and the text path produce the same seed blocks and the same FASTA."""
import io
import os

import numpy as np
import pytest

import dazz_writer as W
from falcon_b200 import consensus, synth
from helpers import emu_engine


def make_case(tmp_path, seed=7, genome=16000, n_reads=36, read_len=2600, with_long=False):
    rng = np.random.default_rng(seed)
    g = synth.random_codes(genome, rng)
    reads, meta = [], []
    for i in range(n_reads):
        ln = int(rng.integers(read_len // 3, read_len * 3 // 2))
        st = int(rng.integers(0, genome - ln))
        strand = int(rng.integers(0, 2))
        x = synth.add_errors(g[st:st + ln], rng, 0.05, 0.03, 0.01)
        s = synth.codes_to_bytes(x)
        reads.append(W.revcomp(s) if strand else s)
        meta.append((st, st + ln, strand))
    if with_long:                                       # a read of more than 100000 bases: cut to 99999 (consensus.py:178-179)
        st, ln = 0, genome
        big = synth.codes_to_bytes(np.concatenate([g, synth.random_codes(100010 - genome, rng)]))
        reads.append(big); meta.append((0, genome, 0))
    overlaps = []
    for a in range(len(reads)):
        sa, ea, stra = meta[a]
        for b in range(len(reads)):
            if a == b:
                continue
            sb, eb, strb = meta[b]
            lo, hi = max(sa, sb), min(ea, eb)
            if hi - lo < 600:
                continue
            alen, blen = len(reads[a]), len(reads[b])
            # coordinates in A's orientation
            if stra == 0:
                d0, d1 = sb - sa, eb - ea
            else:
                d0, d1 = ea - eb, sa - sb
            abpos, bbpos = (0, min(blen - 1, -d0)) if d0 <= 0 else (min(alen - 1, d0), 0)
            aepos, bepos = (alen, max(1, blen - d1)) if d1 >= 0 else (max(1, alen + d1), blen)
            kind = int(rng.integers(0, 10))
            if kind == 0:                                # a local alignment: must be skipped by -o
                abpos, bbpos = max(1, abpos) + 5, max(1, bbpos) + 7
            o = dict(aread=a, bread=b, comp=stra != strb, abpos=abpos, aepos=aepos, bbpos=bbpos, bepos=bepos)
            overlaps.append(o)
            if kind == 1:                                # the same B read twice: the parser keeps the first
                overlaps.append(dict(o))
    overlaps.sort(key=lambda o: (o["aread"], o["bread"]))
    d = str(tmp_path)
    db = W.write_db(d, "reads", reads)
    las = os.path.join(d, "reads.1.las")
    W.write_las(las, overlaps)
    return reads, overlaps, db, las


def run_both(engine, reads, overlaps, db, las, extra, cutoff):
    argv = ["consensus", "--output-multi", "--min-cov", "2", "--min-n-read", "3", "--min-cov-aln", "0"] + extra
    txt = io.StringIO()
    consensus.run(consensus.parse_args(argv), stdin=io.BytesIO(W.la4falcon_text(reads, overlaps, cutoff)), stdout=txt, engine=engine)
    bin_ = io.StringIO()
    consensus.run(consensus.parse_args(argv + ["--db", db, "--las", las, "-H", str(cutoff)]), stdout=bin_, engine=engine)
    return txt.getvalue(), bin_.getvalue()


def test_db_las_path_equals_text_path_emu(tmp_path):
    e = emu_engine()
    reads, overlaps, db, las = make_case(tmp_path)
    for extra, cutoff in (([], 0), (["--max-n-read", "6"], 2500), (["--min-len-aln", "1800", "--max-cov-aln", "2"], 0),
                          (["--trim", "--trim-size", "20"], 0)):
        t, b = run_both(e, reads, overlaps, db, las, extra + ["--batch-blocks", "7"], cutoff)
        assert t == b and t.count(">") > 3, (extra, cutoff)


def test_repacked_store_matches_ascii_upload_emu(tmp_path):
    """k_repack_bps: pool entry 2r / 2r+1 = read r / its reverse complement, bit for bit what k_pack makes of the ASCII."""
    import ctypes as C
    from falcon_b200.binding import DazzDB
    e = emu_engine()
    reads, _, db, _ = make_case(tmp_path, n_reads=9, read_len=700)
    d = DazzDB(db, e._lib)
    d.upload(e)
    assert e.n_reads == 2 * len(reads)
    ptr, n_words, woff = e.pool_device()
    got = np.ctypeslib.as_array((C.c_uint32 * n_words).from_address(ptr)).copy()        # emulator: device memory is host memory
    d.close()
    both = [x for r in reads for x in (r, W.revcomp(r))]
    e.upload_pool(both)
    ptr2, n_words2, woff2 = e.pool_device()
    want = np.ctypeslib.as_array((C.c_uint32 * n_words2).from_address(ptr2)).copy()
    assert n_words == n_words2 and (woff == woff2).all() and (got == want).all()


def test_bad_files_are_reported(tmp_path):
    from falcon_b200.binding import DazzDB, EngineError
    e = emu_engine()
    with pytest.raises(EngineError):
        DazzDB(str(tmp_path / "missing.db"), e._lib)
    reads, overlaps, db, las = make_case(tmp_path, n_reads=12, read_len=2000)
    assert len(overlaps) > 2
    idx = os.path.join(str(tmp_path), ".reads.idx")
    blob = open(idx, "rb").read()
    open(idx, "wb").write(blob[:-8])
    with pytest.raises(EngineError):
        DazzDB(db, e._lib)
    open(idx, "wb").write(blob)
    d = DazzDB(db, e._lib)
    head = open(las, "rb").read()[:50]
    open(las, "wb").write(head)
    d.open_las(las)
    with pytest.raises(EngineError):
        d.take(0, 1, 0, 500, 0, 0, 100, 1 << 20)
    d.close()


@pytest.mark.gpu
def test_db_las_path_equals_text_path_gpu(tmp_path, engine):
    reads, overlaps, db, las = make_case(tmp_path, seed=11, genome=60000, n_reads=150, read_len=5000, with_long=True)
    for extra, cutoff in (([], 0), (["--max-n-read", "20"], 4000), (["--trim"], 0)):
        t, b = run_both(engine, reads, overlaps, db, las, extra, cutoff)
        assert t == b and t.count(">") > 10, (extra, cutoff)
