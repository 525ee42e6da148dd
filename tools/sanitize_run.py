"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from falcon_b200 import synth  # noqa: E402
from falcon_b200.binding import Engine, lib  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

orc = Oracle()
eng = Engine(0)
ok = True
for params, min_cov in ((dict(genome_size=30000, read_len=2500, coverage=20, seed=3, n_blocks=5), 3),
                        (dict(genome_size=20000, read_len=2000, coverage=15, seed=4, n_blocks=3, len_sigma=0.5), 2)):
    S = synth.make_set(**params)
    eng.upload_pool(S.pool)
    got = eng.consensus_blocks([b.tolist() for b in S.blocks], min_cov, 0.70)
    for bi in range(len(S.blocks)):
        ok &= got[bi] == orc.generate_consensus(S.block_seqs(bi), min_cov, 0.70)
# long insertions -> generic consensus path; staged DP; legacy align
rng = np.random.default_rng(2)
g = synth.random_codes(4000, rng)
seed = synth.codes_to_bytes(g)
reads = []
for r in range(8):
    x = synth.add_errors(g, rng, 0.03, 0.02, 0.01)
    reads.append(synth.codes_to_bytes(np.concatenate([x[:1500], synth.random_codes(12, rng), x[1500:]])))
seqs = [seed, seed] + reads
ok &= eng.generate_consensus(seqs, 2, 0.70) == orc.generate_consensus(seqs, 2, 0.70)
for variant in (2, 1, 3):                       # TMA-staged k_dp, plain k_dp (+ k_traceback_walk), back to k_dp3
    eng.set_option("dp_variant", variant)
    ok &= eng.generate_consensus(seqs, 2, 0.70) == orc.generate_consensus(seqs, 2, 0.70)
# device-side --trim (k_trim_range, k_subreads) and the batched distance-only align
S = synth.make_set(30000, 3000, 14, seed=21, n_blocks=2)
eng.upload_pool(S.pool)
boff = np.zeros(len(S.blocks) + 1, dtype=np.uint32)
np.cumsum([len(b) for b in S.blocks], out=boff[1:])
noff, nids = eng.trim_blocks_raw(boff, np.concatenate(S.blocks).astype(np.uint32))
eng.consensus_blocks([nids[int(noff[b]):int(noff[b + 1])].tolist() for b in range(len(S.blocks))], 4, 0.70)
eng.pool_truncate(len(S.pool))
res = eng.align_pairs([2, 3], [0, 0], None, 1500)
ok &= int(res[0][0]) == orc.align(S.pool[2], S.pool[0], 1500)["aln_str_size"]
L = lib()
p = L.align(reads[0], len(reads[0]), seed, len(seed), 150, 1)
ok &= p[0].aln_str_size == orc.align(reads[0], seed)["aln_str_size"]
L.free_alignment(p)
print("SANITIZE RUN", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
