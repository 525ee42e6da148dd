"""CPU: the C-ABI shared library loads and exports every symbol include/falcon_b200.h declares;
struct mirrors have the layouts of the reference's (falcon_kit/falcon_kit.py:19-41,86-106).
No compute calls here (no GPU in the authoring container)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "falcon_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_declares_the_reference_surface():
    fns = declared_functions()
    # the 19 symbols falcon_kit/falcon_kit.py binds at import
    legacy = ["allocate_kmer_lookup", "init_kmer_lookup", "free_kmer_lookup", "allocate_seq", "init_seq_array",
              "free_seq_array", "allocate_seq_addr", "free_seq_addr_array", "add_sequence", "mask_k_mer",
              "find_kmer_pos_for_seq", "free_kmer_match", "find_best_aln_range", "find_best_aln_range2",
              "free_aln_range", "align", "free_alignment", "generate_consensus", "free_consensus_data"]
    for n in legacy:
        assert n in fns, n
    for n in ("fcx_create", "fcx_destroy", "fcx_pool_upload", "fcx_consensus_blocks", "fcx_last_error"):
        assert n in fns, n


def test_library_exports_every_declared_symbol():
    from falcon_b200 import binding
    lib = binding.lib()
    for n in declared_functions():
        assert hasattr(lib, n), "libfalcon_b200.so does not export %s" % n


def test_struct_layouts_match_reference_mirrors():
    from falcon_b200 import binding as b
    assert C.sizeof(b.KmerLookup) == 12
    assert C.sizeof(b.AlnRange) == 24 and b.AlnRange.score.offset == 16
    assert C.sizeof(b.KmerMatch) == 24
    assert C.sizeof(b.ConsensusData) == 16
    assert C.sizeof(b.Alignment) == 40 and b.Alignment.q_aln_str.offset == 24
    assert C.sizeof(b.PairInfo) == 14 * 4


def test_no_gpu_fails_loudly_or_creates_engine():
    """Without a device the engine must refuse (no CPU fallback); with one it must come up."""
    from falcon_b200 import binding
    lib = binding.lib()
    h = C.c_void_p()
    rc = lib.fcx_create(0, C.byref(h))
    if rc != 0:
        msg = lib.fcx_last_error(None).decode()
        assert "no CPU path" in msg or "CUDA" in msg
    else:
        lib.fcx_destroy(h)


def test_host_kmer_helpers_match_oracle(oracle):
    """The legacy k-mer helper symbols are host code operating on ABI-visible host structures."""
    import numpy as np
    from falcon_b200 import binding, synth
    lib = binding.lib()
    rng = np.random.default_rng(4)
    g = synth.random_codes(5000, rng)
    seed = synth.codes_to_bytes(g)
    read = synth.codes_to_bytes(synth.add_errors(g[500:4500], rng))
    K = 8
    lk = lib.allocate_kmer_lookup(1 << (2 * K))
    sa = lib.allocate_seq(len(seed))
    sda = lib.allocate_seq_addr(len(seed))
    lib.add_sequence(0, K, seed, len(seed), sda, sa, lk)
    km = lib.find_kmer_pos_for_seq(read, len(read), K, sda, lk)
    ar = lib.find_best_aln_range(km, K, K * 6, 5)
    want = oracle.kmer_range(read, seed)
    assert (km[0].count, ar[0].s1, ar[0].e1, ar[0].s2, ar[0].e2, ar[0].score) == \
        (want.n_match, want.s1, want.e1, want.s2, want.e2, want.score)
    lib.free_aln_range(ar)
    lib.free_kmer_match(km)
    lib.free_seq_addr_array(sda)
    lib.free_seq_array(sa)
    lib.free_kmer_lookup(lk)
