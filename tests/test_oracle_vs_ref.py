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
