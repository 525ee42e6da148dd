"""Shared test helpers: an oracle-backed stand-in engine for CPU tests of the host logic."""
import io
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleEngine:
    """Same surface as falcon_b200.binding.Engine, arithmetic by the CPU oracle (tests only)."""

    def __init__(self, oracle):
        self.oracle = oracle
        self.pool = []

    def upload_pool(self, reads):
        self.pool = list(reads)

    def consensus_blocks(self, blocks, min_cov, min_idt, K=8):
        return [self.oracle.generate_consensus([self.pool[i] for i in b], min_cov, min_idt, K) for b in blocks]


def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def run_cli(argv, stdin_bytes, engine):
    from falcon_b200 import consensus
    args = consensus.parse_args(["consensus"] + list(argv))
    out = io.StringIO()
    consensus.run(args, stdin=io.BytesIO(stdin_bytes), stdout=out, engine=engine)
    return out.getvalue().encode()
