"""GPU: parity of the CUDA path (through the C ABI) with the CPU oracle and with the golden output
of the unmodified reference CLI.  Bit-exact: this is integer / byte work."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from falcon_b200 import synth
from helpers import GOLDEN, golden_cases, run_cli

pytestmark = pytest.mark.gpu

FIELDS = ("n_match", "s1", "e1", "s2", "e2", "passed_filter", "aligned", "dist", "aln_size", "q_e", "t_e",
          "accepted", "n_tags", "trace_cells")
CASES = golden_cases()


def _check_set(engine, oracle, S, min_cov, min_idt=0.70):
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([b.tolist() for b in S.blocks], min_cov, min_idt)
    info = engine.pair_info()
    p = 0
    for bi in range(len(S.blocks)):
        seqs = S.block_seqs(bi)
        want, oinfo = oracle.generate_consensus(seqs, min_cov, min_idt, want_info=True)
        for j in range(1, len(seqs)):
            g, o = info[p], oinfo[j]
            for f in FIELDS:
                if f in ("dist", "aln_size", "q_e", "t_e") and not o.aligned:
                    continue
                assert getattr(g, f) == getattr(o, f), "block %d pair %d field %s" % (bi, j, f)
            p += 1
        assert got[bi] == want, "block %d consensus differs" % bi
    assert p == len(info)


@pytest.mark.parametrize("params", [
    dict(genome_size=60000, read_len=5000, coverage=30, seed=1, n_blocks=6),
    dict(genome_size=40000, read_len=3000, coverage=20, seed=2, n_blocks=4, len_sigma=0.4),
    dict(genome_size=120000, read_len=15000, coverage=25, seed=3, n_blocks=3),
    dict(genome_size=50000, read_len=4000, coverage=60, seed=4, n_blocks=3, max_n_read=40),
    dict(genome_size=50000, read_len=4000, coverage=15, seed=5, n_blocks=4, p_ins=0.05, p_del=0.05, p_sub=0.05),
])
def test_stage_parity_with_oracle(engine, oracle, params):
    _check_set(engine, oracle, synth.make_set(**params), min_cov=4)


def test_min_cov_and_idt_variants(engine, oracle):
    S = synth.make_set(40000, 4000, 25, seed=6, n_blocks=3)
    for min_cov, min_idt in ((0, 0.70), (1, 0.80), (8, 0.75), (200, 0.70)):
        _check_set(engine, oracle, S, min_cov, min_idt)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cli_matches_reference_golden(name, engine):
    case = CASES[name]
    stdin = open(os.path.join(GOLDEN, case["input"]), "rb").read()
    want = open(os.path.join(GOLDEN, name + ".out"), "rb").read()
    got = run_cli(case["args"], stdin, engine)
    assert hashlib.md5(got).hexdigest() == case["out_md5"]
    assert got == want


def test_edge_blocks(engine, oracle):
    rng = np.random.default_rng(1)
    seed = synth.codes_to_bytes(synth.random_codes(3000, rng))
    other = synth.codes_to_bytes(synth.random_codes(3000, rng))
    short = seed[100:106]
    for seqs in ([seed, seed], [seed, other], [seed], [seed, seed, other, seed[500:2500]], [seed, short, seed],
                 [seed[:90], seed[:90]], [seed, seed[:600], seed[2400:], seed]):
        assert engine.generate_consensus(seqs, 0, 0.70) == oracle.generate_consensus(seqs, 0, 0.70)


def test_low_complexity_match_list_overflow(engine, oracle):
    """A shared 400-base homopolymer gives > 16384 k-mer hits for one pair: the range kernel must
    leave its materialised-list fast path and still agree with the reference semantics."""
    rng = np.random.default_rng(8)
    g = synth.random_codes(6000, rng)
    g[2500:2900] = 0
    seed = synth.codes_to_bytes(g)
    reads = [synth.codes_to_bytes(synth.add_errors(g, rng, 0.03, 0.02, 0.01)) for _ in range(5)]
    seqs = [seed, seed] + reads
    engine.upload_pool(seqs)
    got = engine.consensus_blocks([list(range(len(seqs)))], 2, 0.70)[0]
    info = engine.pair_info()
    want, oinfo = oracle.generate_consensus(seqs, 2, 0.70, want_info=True)
    assert max(i.n_match for i in info) > 16384
    for g_, o_ in zip(info, oinfo[1:]):
        assert (g_.n_match, g_.s1, g_.e1, g_.s2, g_.e2) == (o_.n_match, o_.s1, o_.e1, o_.s2, o_.e2)
    assert got == want


def test_long_insertions_take_the_generic_consensus_path(engine, oracle):
    """Reads carrying 9..20-base insertions (longer than the dense fast path's 7 and than the 11
    bases kept inline in the pile-up entries) must vote through the generic link-list path and the
    xam_lookup() reconstruction, bit-exactly."""
    rng = np.random.default_rng(23)
    g = synth.random_codes(8000, rng)
    seed = synth.codes_to_bytes(g)
    reads = []
    for r in range(14):
        x = synth.add_errors(g, rng, 0.03, 0.02, 0.01)
        parts, pos = [], 0
        for cut in sorted(rng.integers(300, len(x) - 300, size=6)):
            parts.append(x[pos:cut]); parts.append(synth.random_codes(int(rng.integers(9, 21)), rng)); pos = cut
        parts.append(x[pos:])
        reads.append(synth.codes_to_bytes(np.concatenate(parts)))
    # two reads share one identical long insertion so that deep columns get real votes
    ins = synth.random_codes(15, rng)
    for r in (3, 4, 5):
        x = np.frombuffer(reads[r], dtype=np.uint8)
        reads[r] = reads[r][:4000] + synth.codes_to_bytes(ins) + reads[r][4000:]
    seqs = [seed, seed] + reads
    for min_cov in (0, 3):
        got = engine.generate_consensus(seqs, min_cov, 0.70)
        assert got == oracle.generate_consensus(seqs, min_cov, 0.70)
    info = engine.pair_info()
    assert sum(i.accepted for i in info) >= 10


def test_rejects_non_acgt(engine):
    from falcon_b200.binding import EngineError
    with pytest.raises(EngineError):
        engine.upload_pool([b"ACGTNACGT", b"ACGT"])
    with pytest.raises(EngineError):
        engine.upload_pool([b"acgt"])


def test_legacy_generate_consensus_symbol(oracle):
    """The reference's own call: falcon.generate_consensus(char**, n, min_cov, K, min_idt)
    (consensus.py:110-117), including the eqv array."""
    from falcon_b200 import binding
    lib = binding.lib()
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


def test_waves_do_not_change_results(oracle):
    """Splitting a call into many waves (tiny wave limits) must give identical output."""
    from falcon_b200.binding import Engine
    S = synth.make_set(60000, 4000, 25, seed=12, n_blocks=9)
    e1 = Engine(0)
    e1.upload_pool(S.pool)
    a = e1.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)
    os.environ["FCX_WAVE_BLOCKS"] = "2"
    try:
        e2 = Engine(0)
    finally:
        del os.environ["FCX_WAVE_BLOCKS"]
    e2.upload_pool(S.pool)
    b = e2.consensus_blocks([x.tolist() for x in S.blocks], 4, 0.70)
    assert a == b
    assert e2.stats()["waves"] >= 4
    for bi in (0, 8):
        assert a[bi] == oracle.generate_consensus(S.block_seqs(bi), 4, 0.70)


def test_full_size_properties(engine):
    """At BASELINE size (15 kb reads, 50x) the oracle is too slow for the whole set; check
    size-independent properties: determinism, order independence of blocks, and that the consensus
    of each block is close to the seed's error-free template length."""
    S = synth.make_set(400000, 15000, 50, seed=20, n_blocks=24, block_stride=40)
    engine.upload_pool(S.pool)
    blocks = [b.tolist() for b in S.blocks]
    a = engine.consensus_blocks(blocks, 4, 0.70)
    st = engine.stats()
    b = engine.consensus_blocks(blocks, 4, 0.70)
    assert a == b                                     # deterministic
    perm = list(reversed(range(len(blocks))))
    c = engine.consensus_blocks([blocks[i] for i in perm], 4, 0.70)
    assert [c[perm.index(i)] for i in range(len(blocks))] == a   # blocks are independent
    assert st["accepted"] > 0.9 * st["dp_pairs"]
    import re
    for cns in a:
        runs = re.findall(b"[ACGT]+", cns)
        assert runs and max(len(r) for r in runs) > 9000   # a 15 kb seed corrects to a long p-read


def test_full_size_block_vs_oracle(engine, oracle):
    S = synth.make_set(200000, 15000, 50, seed=21, n_blocks=2, block_stride=300)
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)
    for bi in range(len(S.blocks)):
        assert got[bi] == oracle.generate_consensus(S.block_seqs(bi), 4, 0.70)


def test_legacy_align_symbol(oracle):
    """DWA.align (falcon_kit/falcon_kit.py:111-114) on the GPU: band 150 / 100 / 1500, with and
    without alignment strings, aligned and band-failure cases."""
    from falcon_b200 import binding
    lib = binding.lib()
    rng = np.random.default_rng(17)
    g = synth.random_codes(7000, rng)
    a = synth.codes_to_bytes(synth.add_errors(g[200:6800], rng))
    b = synth.codes_to_bytes(synth.add_errors(g[300:6900], rng))
    unrelated = synth.codes_to_bytes(synth.random_codes(3000, rng))
    cases = [(a[100:], b, 150, 1), (a[100:], b, 100, 0), (a[100:], b, 1500, 1), (a[:2500], unrelated, 150, 1),
             (a[:300], a[:300], 150, 1)]
    for q, t, band, want_str in cases:
        p = lib.align(q, len(q), t, len(t), band, want_str)
        got = dict(n=p[0].aln_str_size, dist=p[0].dist, qe=p[0].aln_q_e, te=p[0].aln_t_e,
                   qs=p[0].aln_q_s, ts=p[0].aln_t_s)
        if want_str and got["n"] > 0:
            got["qa"] = C.string_at(p[0].q_aln_str, got["n"]); got["ta"] = C.string_at(p[0].t_aln_str, got["n"])
        lib.free_alignment(p)
        o = oracle.align(q, t, band)
        assert (got["n"], got["qs"], got["ts"]) == (o["aln_str_size"], 0, 0)
        if o["aln_str_size"] > 0:
            assert (got["dist"], got["qe"], got["te"]) == (o["dist"], o["q_e"], o["t_e"])
            if want_str:
                assert got["qa"] == o["q_aln"] and got["ta"] == o["t_aln"]


@pytest.mark.parametrize("read_len,genome", [(30000, 150000), (60000, 240000)])
def test_long_reads_vs_oracle(engine, oracle, read_len, genome):
    """BASELINE config 4 (read-length sweep): 30 kb and 60 kb reads, one block each, bit-exact."""
    S = synth.make_set(genome, read_len, 12, seed=31, n_blocks=1, block_stride=3)
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([b.tolist() for b in S.blocks], 2, 0.70)
    for bi in range(len(S.blocks)):
        assert got[bi] == oracle.generate_consensus(S.block_seqs(bi), 2, 0.70)


def test_block_with_more_than_256_pairs(engine, oracle):
    """--max-n-read defaults to 500 (consensus.py:233): rows beyond the 8 register-resident chunks
    of the consensus kernel are read straight from the pile-up matrix."""
    S = synth.make_set(30000, 2000, 330, seed=41, n_blocks=1, block_stride=250, max_n_read=500, min_ovl=200)
    assert len(S.blocks[0]) > 300
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([S.blocks[0].tolist()], 6, 0.70)[0]
    assert got == oracle.generate_consensus(S.block_seqs(0), 6, 0.70)


def test_wave_split_on_out_of_memory(oracle):
    """A wave that does not fit is split and retried (simulated with the engine's test hook); the
    merged output must be unchanged."""
    from falcon_b200.binding import Engine
    S = synth.make_set(60000, 4000, 25, seed=12, n_blocks=9)
    e = Engine(0)
    e.upload_pool(S.pool)
    blocks = [b.tolist() for b in S.blocks]
    a = e.consensus_blocks(blocks, 4, 0.70)
    e.set_option("debug_split_above", 2)
    b = e.consensus_blocks(blocks, 4, 0.70)
    assert a == b and e.stats()["waves"] >= 5
    assert len(e.pair_info()) == S.n_pairs


def test_unaligned_and_rejected_pairs(engine, oracle):
    """Pairs that pass the k-mer filter but (a) blow the DP band (a 200-base insertion) or (b) align
    with identity below min_idt (their DP runs past the trace bound kept for acceptable pairs)
    must be reported exactly like the reference and must not disturb the block's consensus."""
    rng = np.random.default_rng(29)
    g = synth.random_codes(7000, rng)
    seed = synth.codes_to_bytes(g)
    good = [synth.codes_to_bytes(synth.add_errors(g, rng)) for _ in range(8)]
    band = synth.codes_to_bytes(np.concatenate([g[:3500], synth.random_codes(200, rng), g[3500:]]))
    noisy = [synth.codes_to_bytes(synth.add_errors(g, rng, 0.20, 0.12, 0.06)) for _ in range(4)]
    seqs = [seed, seed] + good[:4] + [band] + noisy + good[4:]
    saw_unaligned = saw_rejected = False
    for min_idt in (0.70, 0.80, 0.55):
        engine.upload_pool(seqs)
        got = engine.consensus_blocks([list(range(len(seqs)))], 3, min_idt)[0]
        info = engine.pair_info()
        want, oinfo = oracle.generate_consensus(seqs, 3, min_idt, want_info=True)
        for j, (g_, o_) in enumerate(zip(info, oinfo[1:])):
            assert (g_.passed_filter, g_.aligned, g_.accepted) == (o_.passed_filter, o_.aligned, o_.accepted), j
            if o_.aligned:
                assert (g_.dist, g_.aln_size, g_.q_e, g_.t_e, g_.trace_cells) == (o_.dist, o_.aln_size, o_.q_e, o_.t_e, o_.trace_cells), j
        assert got == want
        saw_unaligned |= any(o.passed_filter and not o.aligned for o in oinfo[1:])
        saw_rejected |= any(o.aligned and not o.accepted for o in oinfo[1:])
    assert saw_unaligned and saw_rejected


def test_deep_coverage_default_error_model(engine, oracle):
    """~500 reads over a 3 kb seed at 15 % error: many live insertion columns per position (the
    consensus record store scales with coverage); must stay exact."""
    S = synth.make_set(30000, 3000, 480, seed=43, n_blocks=2, block_stride=900, max_n_read=500, min_ovl=500)
    assert max(len(b) for b in S.blocks) > 400
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([b.tolist() for b in S.blocks], 6, 0.70)
    for bi in range(len(S.blocks)):
        assert got[bi] == oracle.generate_consensus(S.block_seqs(bi), 6, 0.70)


def _check_trim(engine, ref, S, edge_tolerance=1000, trim_size=50, max_n_read=500, max_cov_aln=0):
    """fcx_trim_blocks (device chaining + sub-read cutting) vs the reference's get_consensus_with_trim
    logic driven by the compiled reference (tests/ref_host.py): same reads, same order, same bytes --
    checked through the consensus of the trimmed blocks, which must also equal the reference's."""
    import ref_host
    engine.upload_pool(S.pool)
    n0 = engine.n_reads
    block_off = np.zeros(len(S.blocks) + 1, dtype=np.uint32)
    np.cumsum([len(b) for b in S.blocks], out=block_off[1:])
    ids = np.concatenate(S.blocks).astype(np.uint32)
    new_off, new_ids = engine.trim_blocks_raw(block_off, ids, edge_tolerance, trim_size, max_n_read, max_cov_aln)
    cfg = (4, 8, max_n_read, 0.7, edge_tolerance, trim_size, 0, max_cov_aln)
    want_blocks = [ref_host.trim_block(ref, S.block_seqs(bi), cfg) for bi in range(len(S.blocks))]
    for bi, wb in enumerate(want_blocks):
        assert int(new_off[bi + 1] - new_off[bi]) == len(wb), "block %d: number of trimmed reads" % bi
        assert int(new_ids[int(new_off[bi])]) == int(S.blocks[bi][0])
    got = engine.consensus_blocks([new_ids[int(new_off[b]):int(new_off[b + 1])].tolist() for b in range(len(S.blocks))], 4, 0.70)
    for bi, wb in enumerate(want_blocks):
        assert got[bi] == ref.generate_consensus(wb, 4, 0.70), "block %d: consensus of the trimmed block" % bi
    assert sum(len(w) - 1 for w in want_blocks) > 0
    engine.pool_truncate(n0)
    assert engine.consensus_blocks([S.blocks[0].tolist()], 4, 0.70)[0] == ref.generate_consensus(S.block_seqs(0), 4, 0.70)


def test_device_trim_vs_reference(engine, ref):
    _check_trim(engine, ref, synth.make_set(60000, 5000, 25, seed=21, n_blocks=4))
    _check_trim(engine, ref, synth.make_set(120000, 15000, 20, seed=22, n_blocks=3), edge_tolerance=600, trim_size=120)
    _check_trim(engine, ref, synth.make_set(40000, 3000, 40, seed=23, n_blocks=3, len_sigma=0.4), max_n_read=10, max_cov_aln=3)


def test_multi_engine_shards_one_data_set(engine, oracle):
    """fcx_multi_*: one read store on every device (peer copies), seed blocks sharded, results merged
    in seed order.  Uses every visible GPU; on a one-GPU box two engines share device 0."""
    from falcon_b200.binding import MultiEngine, device_count
    n = device_count()
    devs = list(range(n)) if n > 1 else [0, 0]
    S = synth.make_set(80000, 4000, 20, seed=31, n_blocks=9)
    blocks = [b.tolist() for b in S.blocks]
    engine.upload_pool(S.pool)
    single = engine.consensus_blocks(blocks, 4, 0.70)
    m = MultiEngine(devs)
    try:
        m.upload_pool(S.pool)
        multi = m.consensus_blocks(blocks, 4, 0.70)
        assert m.peer_bytes() > 0
        st = m.stats()
        assert st["pairs"] == S.n_pairs
    finally:
        m.close()
    assert multi == single
    for bi in (0, len(blocks) - 1):
        assert multi[bi] == oracle.generate_consensus(S.block_seqs(bi), 4, 0.70)


def test_align_pairs_batched_vs_reference(engine, ref):
    """fcx_align_pairs (graph_to_contig.get_aln_data: DWA.align(.., 1500, 1)) against the compiled reference."""
    rng = np.random.default_rng(41)
    seqs, want, q_ids, t_ids = [], [], [], []
    for i in range(12):
        g = synth.random_codes(int(rng.integers(800, 6000)), rng)
        a = synth.codes_to_bytes(g)
        b = synth.codes_to_bytes(synth.add_errors(g, rng, 0.02 * (i % 4), 0.02 * (i % 3), 0.01))
        q_ids.append(len(seqs)); seqs.append(b)
        t_ids.append(len(seqs)); seqs.append(a)
        r = ref.align(b, a, 1500)
        want.append((r["aln_str_size"], r["dist"], r["q_e"], r["t_e"]))
    engine.upload_pool(seqs)
    got = engine.align_pairs(q_ids, t_ids, None, 1500)
    for i, w in enumerate(want):
        assert tuple(int(x) for x in got[i]) == (w if w[0] else (0, 0, 0, 0)), i


def test_capacity_retry_paths(oracle):
    """Vote overflow arena and consensus record arena that start far too small: the wave is retried with
    larger arenas instead of failing the call (the reference reallocs)."""
    from falcon_b200.binding import Engine
    e = Engine(0)
    try:
        e.set_option("debug_tiny_capacity", 1)
        S = synth.make_set(40000, 3000, 30, seed=51, n_blocks=3)
        e.upload_pool(S.pool)
        got = e.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)
        for bi in range(len(S.blocks)):
            assert got[bi] == oracle.generate_consensus(S.block_seqs(bi), 4, 0.70)
    finally:
        e.close()


def test_many_full_size_blocks_vs_compiled_reference(engine, ref):
    """64 full-size blocks (15 kb reads, 50x, BASELINE config 2 geometry) against oracle/_ref/falcon.so --
    the unmodified reference C -- not the restatement."""
    S = synth.make_set(4_600_000, 15000, 50, n_blocks=64, max_n_read=200, block_stride=239)
    engine.upload_pool(S.pool)
    got = engine.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)
    from concurrent.futures import ThreadPoolExecutor
    import os as _os
    # the reference is not re-entrant (static msa_array): fork workers, as falcon_kit does
    import multiprocessing as mp
    jobs = [(S.block_seqs(bi), 4, 0.70) for bi in range(len(S.blocks))]
    with mp.get_context("fork").Pool(min(16, len(_os.sched_getaffinity(0)))) as pool:
        want = pool.starmap(_ref_block, jobs)
    for bi, w in enumerate(want):
        assert got[bi] == w, "block %d" % bi


def _ref_block(seqs, min_cov, min_idt):
    from oracle.oracle import Ref
    return Ref().generate_consensus(seqs, min_cov, min_idt)


def test_rare_paths_also_covered_on_the_gpu(engine, oracle):
    """The bodies of the emulator-only tests (wide DP bands, > 15 and > 32 links per position, deep
    pile-up of > 256 pairs) run against the real device."""
    import test_emu_parity as E
    E.test_wide_bands_stay_exact(engine, oracle)
    E.test_deep_coverage_overflows_the_position_slots(engine, oracle)
    E.test_deep_noisy_pileup_takes_the_chunked_consensus_path(engine, oracle)


def test_round1_dp_variants_still_agree(oracle):
    """dp_variant 1 (k_dp) and 2 (k_dp with TMA-staged spans) + k_traceback_walk."""
    from falcon_b200.binding import Engine
    S = synth.make_set(40000, 4000, 20, seed=61, n_blocks=3)
    for variant in (1, 2):
        e = Engine(0)
        try:
            e.set_option("dp_variant", variant)
            _check_set(e, oracle, S, min_cov=4)
        finally:
            e.close()
