"""CPU: pin the C restatement (oracle/fc_oracle.c) against the compiled, unmodified reference
(oracle/_ref/falcon.so) stage by stage.  Skipped when the prebuilt reference .so is absent."""
import numpy as np
import pytest

from falcon_b200 import synth


def _fa(path):
    return b"".join(l.strip().encode() for l in open(path) if not l.startswith(">"))


@pytest.fixture(scope="module")
def sets():
    return [synth.make_set(40000, 4000, 25, seed=3, n_blocks=4),
            synth.make_set(30000, 2500, 20, seed=5, n_blocks=4, len_sigma=0.5)]


def test_ranges_match_reference(oracle, ref, sets):
    for S in sets:
        seqs = S.block_seqs(0)
        for r in seqs[1:12]:
            want = ref.kmer_range(r, seqs[0])
            got = oracle.kmer_range(r, seqs[0])
            assert (got.n_match, got.s1, got.e1, got.s2, got.e2, got.score) == want


def test_align_strings_match_reference(oracle, ref, sets):
    S = sets[0]
    seqs = S.block_seqs(1)
    n = 0
    for r in seqs[1:]:
        g = oracle.kmer_range(r, seqs[0])
        if g.e1 - g.s1 < 100:
            continue
        q, t = r[g.s1:g.e1], seqs[0][g.s2:g.e2]
        a, b = ref.align(q, t), oracle.align(q, t)
        for k in ("aln_str_size", "dist", "q_e", "t_e", "q_aln", "t_aln"):
            assert a[k] == b[k]
        n += 1
    assert n > 5


def test_align_band_failure_and_tiny(oracle, ref):
    rng = np.random.default_rng(0)
    a = synth.codes_to_bytes(synth.random_codes(3000, rng))
    b = synth.codes_to_bytes(synth.random_codes(3000, rng))
    ra, oa = ref.align(a, b), oracle.align(a, b)        # unrelated sequences: no alignment
    assert ra["aln_str_size"] == oa["aln_str_size"]
    assert ra["dist"] == oa["dist"]
    ra, oa = ref.align(a[:200], a[:200]), oracle.align(a[:200], a[:200])
    assert (ra["aln_str_size"], ra["dist"]) == (oa["aln_str_size"], oa["dist"]) == (200, 0)


def test_generate_consensus_matches_reference(oracle, ref, sets):
    for S in sets:
        for bi in range(len(S.blocks)):
            seqs = S.block_seqs(bi)
            for min_cov in (0, 4):
                want, weqv = ref.generate_consensus(seqs, min_cov, 0.70, want_eqv=True)
                got, geqv = oracle.generate_consensus(seqs, min_cov, 0.70, want_eqv=True)
                assert got == want
                assert geqv == weqv


def test_edge_blocks(oracle, ref):
    rng = np.random.default_rng(1)
    seed = synth.codes_to_bytes(synth.random_codes(3000, rng))
    other = synth.codes_to_bytes(synth.random_codes(3000, rng))
    # seed only with itself; seed with an unrelated read; n_seq == 1 (nothing aligned -> "")
    for seqs in ([seed, seed], [seed, other], [seed], [seed, seed, other, seed[500:2500]]):
        assert oracle.generate_consensus(seqs, 0, 0.70) == ref.generate_consensus(seqs, 0, 0.70)


def test_long_insertion_delta_wrap(oracle, ref):
    """A read with a 300-base insertion drives delta past the uint8 bookkeeping of the reference's
    column store (falcon.c:363-368, 205-218); the restatement must follow it."""
    rng = np.random.default_rng(2)
    g = synth.random_codes(6000, rng)
    seed = synth.codes_to_bytes(g)
    reads = [seed]
    for k in range(6):
        reads.append(synth.codes_to_bytes(synth.add_errors(g, rng, 0.02, 0.01, 0.005)))
    ins = synth.random_codes(300, rng)
    reads.append(synth.codes_to_bytes(np.concatenate([g[:3000], ins, g[3000:]])))
    reads.append(synth.codes_to_bytes(synth.add_errors(g, rng, 0.02, 0.01, 0.005)))
    seqs = [seed] + reads
    assert oracle.generate_consensus(seqs, 2, 0.70) == ref.generate_consensus(seqs, 2, 0.70)


def test_randomised_blocks_match_reference(oracle, ref):
    """A sweep over error mixes, read-length spreads, coverages and min_cov/min_idt settings: the
    restatement must track the compiled reference on every block (consensus + eqv)."""
    rng = np.random.default_rng(77)
    n = 0
    for trial in range(10):
        p_ins, p_del, p_sub = (float(x) for x in rng.uniform([0.02, 0.01, 0.0], [0.12, 0.08, 0.05]))
        S = synth.make_set(int(rng.integers(20000, 50000)), int(rng.integers(1500, 5000)), float(rng.uniform(8, 35)),
                           seed=int(rng.integers(1, 10**6)), n_blocks=3, len_sigma=float(rng.choice([0.0, 0.3, 0.6])),
                           p_ins=p_ins, p_del=p_del, p_sub=p_sub, block_stride=int(rng.integers(1, 6)),
                           max_n_read=int(rng.choice([8, 40, 200])))
        min_cov = int(rng.integers(0, 7)); min_idt = float(rng.choice([0.6, 0.7, 0.8, 0.9]))
        for bi in range(len(S.blocks)):
            seqs = S.block_seqs(bi)
            want = ref.generate_consensus(seqs, min_cov, min_idt, want_eqv=True)
            got = oracle.generate_consensus(seqs, min_cov, min_idt, want_eqv=True)
            assert got[0] == want[0] and got[1] == want[1], (trial, bi)
            n += 1
    assert n == 30
