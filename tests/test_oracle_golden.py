"""CPU: the oracle restatement + our host CLI logic reproduce the reference CLI's golden output
(tests/golden/*.out were produced by the unmodified reference, see make_golden.py)."""
import hashlib
import os

import pytest

from helpers import GOLDEN, OracleEngine, golden_cases, run_cli

CASES = golden_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_cli_with_oracle_matches_reference_golden(name, oracle):
    case = CASES[name]
    stdin = open(os.path.join(GOLDEN, case["input"]), "rb").read()
    assert hashlib.md5(stdin).hexdigest() == case["in_md5"]
    want = open(os.path.join(GOLDEN, name + ".out"), "rb").read()
    assert hashlib.md5(want).hexdigest() == case["out_md5"]
    got = run_cli(case["args"], stdin, OracleEngine(oracle))
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
    from falcon_b200 import consensus
    cfg = (4, 8, 500, 0.7, 1000, 50, 0, 0)
    long_seq = b"A" * 100005
    txt = b"s1 ACGT\nr1 AAAA\nr1 CCCC\nbad line here\nr2 GG\n+ +\nx1 TTTT\n* *\ny1 " + long_seq + b"\n+ +\n- -\nz1 ACGT\n+ +\n"
    blocks = list(consensus.get_seq_data(io.BytesIO(txt), cfg, 1, 0))
    assert [sid for _, sid in blocks] == ["s1", "y1"]
    assert blocks[0][0] == [b"ACGT", b"ACGT", b"AAAA", b"GG"]
    assert len(blocks[1][0][0]) == 99999 and len(blocks[1][0]) == 2
