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
eng.set_option("dp_staged", 1)
ok &= eng.generate_consensus(seqs, 2, 0.70) == orc.generate_consensus(seqs, 2, 0.70)
L = lib()
p = L.align(reads[0], len(reads[0]), seed, len(seed), 150, 1)
ok &= p[0].aln_str_size == orc.align(reads[0], seed)["aln_str_size"]
L.free_alignment(p)
print("SANITIZE RUN", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
