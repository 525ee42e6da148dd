"""CPU: the oracle restatement + our host CLI logic reproduce the reference CLI's golden output
(tests/golden/*.out were produced by the unmodified reference, see make_golden.py)."""
import hashlib
import os

import pytest

from helpers import GOLDEN, OracleEngine, golden_cases, run_cli, run_cli_python_host

CASES = golden_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_cli_with_oracle_matches_reference_golden(name, oracle, request):
    case = CASES[name]
    ref = request.getfixturevalue("ref") if "--trim" in case["args"] else None
    stdin = open(os.path.join(GOLDEN, case["input"]), "rb").read()
    assert hashlib.md5(stdin).hexdigest() == case["in_md5"]
    want = open(os.path.join(GOLDEN, name + ".out"), "rb").read()
    assert hashlib.md5(want).hexdigest() == case["out_md5"]
    got = run_cli(case["args"], stdin, OracleEngine(oracle, ref))                  # native parser
    assert got == want
    got = run_cli_python_host(case["args"], stdin, OracleEngine(oracle, ref), ref)  # Python restatement of the host side
    assert got == want


def test_survey_known_answer_md5s():
    # md5s of the reference CLI on the t1/t2 blocks
    known = {"t1t2_a_default_cov0": "b3367d2eaaddfdf2ec37c404e9eefcbb",
             "t1t2_a_default_cov1": "aeda62435c48f7a26c26746c24c0e381",
             "t1t2_a_multi_cov0": "dc56ae1a9409f88ce47e0de6c2efb92b",
             "t1t2_a_full_cov1": "8cd0a52f45a75e678208995a4406fe10",
             "t1t2_b_full_cov1": "391de5f9f6c797a74b8d10226c93b250"}
    for k, v in known.items():
        assert CASES[k]["out_md5"] == v


def test_underscore_flags_accepted():
    from falcon_b200 import consensus
    a = consensus.parse_args(["x", "--output_multi", "--min_idt", "0.8", "--min_cov=3", "--max_n_read", "77"])
    assert a.output_multi and a.min_idt == 0.8 and a.min_cov == 3 and a.max_n_read == 77


def test_parser_block_rules():
    # consensus.py:161-209: 2-token lines only, '+' emits, '*' discards, '-' stops, dup ids dropped,
    # seed appended twice, >100000 cut to 99999
    import io
    import ref_host as consensus
    cfg = (4, 8, 500, 0.7, 1000, 50, 0, 0)
    long_seq = b"A" * 100005
    txt = b"s1 ACGT\nr1 AAAA\nr1 CCCC\nbad line here\nr2 GG\n+ +\nx1 TTTT\n* *\ny1 " + long_seq + b"\n+ +\n- -\nz1 ACGT\n+ +\n"
    blocks = list(consensus.get_seq_data(io.BytesIO(txt), cfg, 1, 0))
    assert [sid for _, sid in blocks] == ["s1", "y1"]
    assert blocks[0][0] == [b"ACGT", b"ACGT", b"AAAA", b"GG"]
    assert len(blocks[1][0][0]) == 99999 and len(blocks[1][0]) == 2


def _python_blocks(txt, min_n_read, min_len_aln, max_n_read, min_cov_aln, max_cov_aln):
    import io
    import ref_host as consensus
    cfg = (4, 8, max_n_read, 0.7, 1000, 50, min_cov_aln, max_cov_aln)
    return [(sid, seqs) for seqs, sid in consensus.get_seq_data(io.BytesIO(txt), cfg, min_n_read, min_len_aln)]


def _native_blocks(txt, min_n_read, min_len_aln, max_n_read, min_cov_aln, max_cov_aln, chunk):
    import ctypes as C
    from falcon_b200.binding import StreamParser
    ps = StreamParser(min_n_read, min_len_aln, max_n_read, min_cov_aln, max_cov_aln)
    out = []
    pos = 0
    while True:
        piece = txt[pos:pos + chunk]
        pos += chunk
        eof = pos >= len(txt)
        n = ps.feed(piece, eof)
        while n > 0:
            ptr, off, boff, rids, ids = ps.take(3, 1 << 20)
            for b, sid in enumerate(ids):
                seqs = [C.string_at(ptr + int(off[r]), int(off[r + 1] - off[r])) for r in rids[int(boff[b]):int(boff[b + 1])]]
                out.append((sid, seqs))
            n = ps.pending()
        if eof or ps.stopped:
            break
    return out


def test_native_parser_matches_python_restatement():
    """fcx_parser_* vs the Python get_seq_data + get_longest_reads on tricky streams, with the
    stream cut at arbitrary chunk boundaries."""
    import random
    rnd = random.Random(5)
    def seq(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))
    lines = []
    for b in range(12):
        n = rnd.randint(0, 9)
        sid = "seed%d" % b
        lines.append("%s %s" % (sid, seq(rnd.randint(5, 400))))
        for r in range(n):
            rid = rnd.choice(["r%d_%d" % (b, r), "r%d_%d" % (b, max(0, r - 1)), sid])   # duplicates
            lines.append("%s  \t %s  " % (rid, seq(rnd.randint(1, 500))))
        if rnd.random() < 0.2:
            lines.append("garbage with four tokens here")
        if rnd.random() < 0.2:
            lines.append("")
        lines.append(rnd.choice(["+ +", "+ +", "+ +", "* *", "+ x"]))
    lines.append("big " + "A" * 100007)
    lines.append("other " + "C" * 100000)
    lines.append("+ +")
    lines.append("- -")
    lines.append("after %s" % seq(50))
    lines.append("+ +")
    txt = ("\n".join(lines) + "\n").encode()
    for params in ((1, 0, 500, 0, 0), (3, 50, 4, 0, 0), (2, 0, 500, 1, 2), (1, 0, 3, 0, 1)):
        want = _python_blocks(txt, *params)
        for chunk in (7, 1000, 1 << 20):
            got = _native_blocks(txt, *params, chunk)
            assert got == want, (params, chunk)
    assert len(_python_blocks(txt, 1, 0, 500, 0, 0)) >= 5


def test_native_parser_parallel_feed_matches_python_model(monkeypatch):
    """Chunks with three or more control lines are parsed block-parallel (worker threads with fresh state);
    the result must not depend on the thread count, the chunking, or where a chunk cuts a line."""
    import random
    rnd = random.Random(11)

    def seq(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))
    lines = []
    for b in range(40):
        sid = "s%d" % b
        lines.append("%s %s" % (sid, seq(rnd.randint(50, 900))))
        for r in range(rnd.randint(0, 12)):
            lines.append("%s %s" % (rnd.choice(["r%d_%d" % (b, r), sid, "r%d_0" % b]), seq(rnd.randint(1, 1200))))
        if rnd.random() < 0.15:
            lines.append("+ one two")                      # three tokens: not a control line
        if rnd.random() < 0.15:
            lines.append("  +\t" + seq(30))                # "+" with a sequence as second token IS a control line
        else:
            lines.append(rnd.choice(["+ +", "+ +", "* *", " + + "]))
    lines += ["tail1 " + seq(400), "tail2 " + seq(300), "+ +", "- -", "after " + seq(100), "+ +"]
    txt = ("\n".join(lines) + "\n").encode()
    want = _python_blocks(txt, 2, 0, 8, 0, 0)
    assert len(want) > 15
    monkeypatch.setenv("FCX_PARSER_PAR_MIN", "0")
    for threads in ("1", "2", "5"):
        monkeypatch.setenv("FCX_PARSER_THREADS", threads)
        for chunk in (997, 20000, 1 << 22):
            assert _native_blocks(txt, 2, 0, 8, 0, 0, chunk) == want, (threads, chunk)
