"""CPU: the CUDA kernel SOURCES compiled against the SIMT emulator (tests/emu) and run through the
same C ABI as on the GPU, checked against the oracle bit for bit.  This is how kernel logic is
validated on the authoring container (no GPU); the `gpu`-marked tests repeat it on the B200.
The emulator library is test infrastructure: falcon_b200.binding never loads it."""
import numpy as np
import pytest

from falcon_b200 import synth
from helpers import emu_engine

import test_gpu_parity as G


@pytest.fixture(scope="module")
def emu():
    return emu_engine()


@pytest.mark.parametrize("params", [
    dict(genome_size=30000, read_len=2500, coverage=20, seed=1, n_blocks=3),
    dict(genome_size=30000, read_len=3000, coverage=14, seed=2, n_blocks=3, len_sigma=0.4),
    dict(genome_size=60000, read_len=8000, coverage=12, seed=3, n_blocks=2),
    dict(genome_size=30000, read_len=3000, coverage=15, seed=5, n_blocks=3, p_ins=0.05, p_del=0.05, p_sub=0.05),
])
def test_stage_parity_with_oracle(emu, oracle, params):
    G._check_set(emu, oracle, synth.make_set(**params), min_cov=4)


def test_min_cov_and_idt_variants(emu, oracle):
    S = synth.make_set(30000, 3000, 16, seed=6, n_blocks=2)
    for min_cov, min_idt in ((0, 0.70), (1, 0.80), (200, 0.70)):
        G._check_set(emu, oracle, S, min_cov, min_idt)


def test_edge_blocks(emu, oracle):
    G.test_edge_blocks(emu, oracle)


def test_low_complexity_match_list_overflow(emu, oracle):
    G.test_low_complexity_match_list_overflow(emu, oracle)


def test_long_insertions_take_the_generic_consensus_path(emu, oracle):
    G.test_long_insertions_take_the_generic_consensus_path(emu, oracle)


def test_unaligned_and_rejected_pairs(emu, oracle):
    """includes a 200-base insertion: the DP band passes 64 cells (wide mode of k_dp3) and then
    exceeds the band limit"""
    G.test_unaligned_and_rejected_pairs(emu, oracle)


def test_wide_bands_stay_exact(emu, oracle):
    """Low-complexity pairs (shared homopolymer / dinucleotide runs) keep many diagonals within the
    band tolerance: bands of 65..151 cells run in k_dp3's shared-memory wide mode."""
    rng = np.random.default_rng(77)
    g = synth.random_codes(5000, rng)
    g[1500:1700] = 1
    g[3000:3300] = np.tile([0, 2], 150)
    seed = synth.codes_to_bytes(g)
    reads = [synth.codes_to_bytes(synth.add_errors(g, rng, 0.05, 0.03, 0.01)) for _ in range(6)]
    seqs = [seed, seed] + reads
    emu.upload_pool(seqs)
    got = emu.consensus_blocks([list(range(len(seqs)))], 2, 0.70)[0]
    info = emu.pair_info()
    want, oinfo = oracle.generate_consensus(seqs, 2, 0.70, want_info=True)
    for j, (g_, o_) in enumerate(zip(info, oinfo[1:])):
        assert (g_.aligned, g_.accepted, g_.trace_cells) == (o_.aligned, o_.accepted, o_.trace_cells), j
        if o_.aligned:
            assert (g_.dist, g_.aln_size, g_.q_e, g_.t_e) == (o_.dist, o_.aln_size, o_.q_e, o_.t_e), j
    assert got == want


@pytest.mark.parametrize("variant", [1])
def test_round1_dp_kernel_still_agrees(oracle, variant):
    e = emu_engine()
    e.set_option("dp_variant", variant)
    G._check_set(e, oracle, synth.make_set(30000, 2500, 14, seed=9, n_blocks=2), min_cov=3)


def test_deep_coverage_overflows_the_position_slots(emu, oracle):
    """~300 reads over a 2 kb seed: far more than 15 distinct links per position, so links spill to
    the overflow arena, and more than 256 pairs per block."""
    S = synth.make_set(20000, 2000, 330, seed=41, n_blocks=1, block_stride=250, max_n_read=500, min_ovl=200)
    assert len(S.blocks[0]) > 250
    emu.upload_pool(S.pool)
    got = emu.consensus_blocks([S.blocks[0].tolist()], 6, 0.70)[0]
    assert got == oracle.generate_consensus(S.block_seqs(0), 6, 0.70)


def test_eqv_and_legacy_symbol(emu, oracle):
    """generate_consensus(char**, ...) of the emulator library, including the eqv array."""
    import ctypes as C
    lib = emu._lib
    S = synth.make_set(30000, 3000, 20, seed=9, n_blocks=2)
    for bi in range(2):
        seqs = S.block_seqs(bi)
        arr = (C.c_char_p * len(seqs))(*seqs)
        p = lib.generate_consensus(arr, len(seqs), 4, 8, 0.70)
        cns = C.string_at(p[0].sequence)
        eqv = [C.cast(p[0].eff_cov, C.POINTER(C.c_int))[i] for i in range(len(cns))]
        lib.free_consensus_data(p)
        want, weqv = oracle.generate_consensus(seqs, 4, 0.70, want_eqv=True)
        assert cns == want
        assert eqv == weqv


def test_waves_and_capacity_retry(oracle):
    """Tiny waves, a simulated out-of-memory split and a forced capacity retry give the same bytes."""
    import os
    S = synth.make_set(40000, 3000, 18, seed=12, n_blocks=7)
    e1 = emu_engine()
    e1.upload_pool(S.pool)
    blocks = [b.tolist() for b in S.blocks]
    a = e1.consensus_blocks(blocks, 4, 0.70)
    e1.set_option("debug_split_above", 2)
    assert e1.consensus_blocks(blocks, 4, 0.70) == a and e1.stats()["waves"] >= 4
    e1.set_option("debug_split_above", 0)
    e1.set_option("debug_tiny_capacity", 1)       # first attempt of every wave overflows its arenas
    assert e1.consensus_blocks(blocks, 4, 0.70) == a
    for bi in (0, 6):
        assert a[bi] == oracle.generate_consensus(S.block_seqs(bi), 4, 0.70)


def test_multi_engine_shards_and_merges_in_order(emu, oracle):
    """fcx_multi_* with 3 (emulated) devices: the read store is assembled from per-device parts,
    blocks are sharded, the merged output equals the single-engine output block for block."""
    from helpers import emu_multi_engine
    S = synth.make_set(40000, 3000, 16, seed=21, n_blocks=8)
    blocks = [b.tolist() for b in S.blocks]
    emu.upload_pool(S.pool)
    single = emu.consensus_blocks(blocks, 3, 0.70)
    info1 = emu.pair_info()
    m = emu_multi_engine(3)
    m.upload_pool(S.pool)
    assert m.peer_bytes() > 0
    multi = m.consensus_blocks(blocks, 3, 0.70)
    assert multi == single
    info3 = m.pair_info()
    assert [(i.s1, i.e1, i.dist, i.accepted) for i in info3] == [(i.s1, i.e1, i.dist, i.accepted) for i in info1]
    assert m.stats()["pairs"] == S.n_pairs
    assert multi[0] == oracle.generate_consensus(S.block_seqs(0), 3, 0.70)
    # fewer blocks than devices, and a single block
    assert m.consensus_blocks(blocks[:2], 3, 0.70) == single[:2]
    assert m.consensus_blocks(blocks[5:6], 3, 0.70) == single[5:6]


def test_cli_streams_and_devices(tmp_path, emu, oracle):
    """The drop-in command with several producers: two --stream IN:OUT pairs served by one process
    (parser threads + one GPU loop), each output in its own input order and equal to the
    single-stream output; and the stdin path equal to the oracle-backed host logic."""
    import io
    from falcon_b200 import consensus
    from helpers import OracleEngine
    S = synth.make_set(30000, 2500, 14, seed=5, n_blocks=6)
    texts = [S.la4falcon_text([0, 1, 2]), S.la4falcon_text([3, 4, 5])]
    want = []
    for t in texts:
        out = io.StringIO()
        a = consensus.parse_args(["consensus", "--output-multi", "--min-cov", "2", "--min-n-read", "1", "--min-cov-aln", "0"])
        consensus.run(a, stdin=io.BytesIO(t), stdout=out, engine=OracleEngine(oracle))
        want.append(out.getvalue())
        assert out.getvalue().count(">") >= 3
    argv = ["consensus", "--output-multi", "--min-cov", "2", "--min-n-read", "1", "--min-cov-aln", "0", "--batch-blocks", "2"]
    for i, t in enumerate(texts):
        (tmp_path / ("in%d" % i)).write_bytes(t)
        argv += ["--stream", "%s:%s" % (tmp_path / ("in%d" % i), tmp_path / ("out%d" % i))]
    consensus.run(consensus.parse_args(argv), engine=emu)
    for i in range(2):
        assert (tmp_path / ("out%d" % i)).read_text() == want[i]


def test_batched_align_pairs(emu, oracle):
    """fcx_align_pairs: the graph_to_contig-style call (band 1500, sub-ranges, sequences longer than
    the 100000-base limit of the consensus path) against the oracle's align()."""
    rng = np.random.default_rng(3)
    g = synth.random_codes(120000, rng)
    a = synth.codes_to_bytes(synth.add_errors(g[:9000], rng, 0.02, 0.01, 0.01))
    b = synth.codes_to_bytes(synth.add_errors(g[200:9500], rng, 0.02, 0.01, 0.01))
    big_a = synth.codes_to_bytes(synth.add_errors(g, rng, 0.005, 0.003, 0.002))      # > 100000 bases
    big_b = synth.codes_to_bytes(synth.add_errors(g, rng, 0.005, 0.003, 0.002))
    unrelated = synth.codes_to_bytes(synth.random_codes(4000, rng))
    pool = [a, b, big_a, big_b, unrelated]
    emu.upload_pool(pool)
    jobs = [(0, 1, (250, 8000, 40, 7900), 1500), (0, 1, (250, 8000, 40, 7900), 150), (0, 4, None, 1500),
            (2, 3, (100000, 110000, 100000, 110050), 1500)]
    for qi, ti, rg, band in jobs:
        got = emu.align_pairs([qi], [ti], None if rg is None else [rg], band)[0]
        q = pool[qi] if rg is None else pool[qi][rg[0]:rg[1]]
        t = pool[ti] if rg is None else pool[ti][rg[2]:rg[3]]
        o = oracle.align(q, t, band)
        assert got[0] == o["aln_str_size"]
        if o["aln_str_size"] > 0:
            assert (got[1], got[2], got[3]) == (o["dist"], o["q_e"], o["t_e"])
    many = emu.align_pairs([0, 0, 1], [1, 1, 0], None, 1500)
    assert (many[0] == many[1]).all() and many[0][0] > 0


def test_device_trim_vs_reference(emu, ref):
    """k_trim_range / k_subreads / fcx_trim_blocks against the compiled reference's --trim logic."""
    G._check_trim(emu, ref, synth.make_set(30000, 3000, 14, seed=21, n_blocks=3))
    G._check_trim(emu, ref, synth.make_set(30000, 2500, 20, seed=23, n_blocks=2, len_sigma=0.4), edge_tolerance=500,
                  trim_size=80, max_n_read=6, max_cov_aln=2)


def test_deep_noisy_pileup_takes_the_chunked_consensus_path(emu, oracle):
    """> 32 distinct links at one seed position (deep coverage, high insertion rate): k_cns_dp's
    chunked slow path and the vote overflow arena."""
    rng = np.random.default_rng(5)
    g = synth.random_codes(700, rng)
    seed = synth.codes_to_bytes(g)
    reads = [synth.codes_to_bytes(synth.add_errors(g, rng, 0.13, 0.11, 0.01)) for _ in range(1500)]   # 6 positions with > 32 links
    seqs = [seed, seed] + reads
    emu.upload_pool(seqs)
    got = emu.consensus_blocks([list(range(len(seqs)))], 4, 0.60)[0]
    assert got == oracle.generate_consensus(seqs, 4, 0.60)


def test_randomised_parameter_sweep(emu, oracle):
    """A few random workloads (read length, coverage, error mix, min_cov, min_idt): tools/emu_stress.py runs
    the long version of this."""
    rng = np.random.default_rng(2026)
    for _ in range(6):
        rl = int(rng.choice([1200, 2500, 4000]))
        S = synth.make_set(genome_size=int(rl * rng.integers(6, 12)), read_len=rl, coverage=float(rng.choice([8, 15, 25])),
                           seed=int(rng.integers(1, 1 << 30)), n_blocks=2, p_ins=float(rng.choice([0.02, 0.09, 0.13])),
                           p_del=float(rng.choice([0.01, 0.045, 0.09])), p_sub=float(rng.choice([0.0, 0.015, 0.04])),
                           len_sigma=float(rng.choice([0.0, 0.35])), max_n_read=int(rng.choice([12, 200])))
        G._check_set(emu, oracle, S, int(rng.choice([0, 1, 4])), float(rng.choice([0.6, 0.7, 0.8])))


def test_query_read_of_exactly_100000_bases(emu, oracle):
    """consensus.py:178-179 cuts reads only when they are LONGER than 100000; a 100000-base query read is
    legal (falcon.c:343 limits the seed only).  The engine accepts it; a 100000-base SEED is refused."""
    from falcon_b200.binding import EngineError
    rng = np.random.default_rng(9)
    g = synth.random_codes(100000, rng)
    seed = synth.codes_to_bytes(g[40000:42500])
    long_read = synth.codes_to_bytes(g)                              # contains the seed region exactly
    reads = [synth.codes_to_bytes(synth.add_errors(g[40000:42500], rng, 0.05, 0.03, 0.01)) for _ in range(5)]
    seqs = [seed, seed, long_read] + reads
    assert len(long_read) == 100000
    emu.upload_pool(seqs)
    got = emu.consensus_blocks([list(range(len(seqs)))], 2, 0.70)[0]
    assert got == oracle.generate_consensus(seqs, 2, 0.70)
    with pytest.raises(EngineError):
        emu.consensus_blocks([[2, 2, 0]], 2, 0.70)
