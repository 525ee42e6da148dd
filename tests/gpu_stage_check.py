"""Stage-by-stage GPU-vs-oracle comparison (developer tool; the pytest suite wraps the same checks).

    python tests/gpu_stage_check.py [--genome 60000 --len 5000 --cov 30 --blocks 6]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from falcon_b200 import synth  # noqa: E402
from falcon_b200.binding import Engine  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

FIELDS = ("n_match", "s1", "e1", "s2", "e2", "passed_filter", "aligned", "dist", "aln_size", "q_e",
          "t_e", "accepted", "n_tags", "trace_cells")


def compare(eng, orc, S, min_cov=4, min_idt=0.70, verbose=True):
    eng.upload_pool(S.pool)
    t0 = time.time()
    out = eng.consensus_blocks([b.tolist() for b in S.blocks], min_cov, min_idt)
    dt = time.time() - t0
    info = eng.pair_info()
    bad_blocks = 0
    bad_pairs = 0
    p = 0
    for bi in range(len(S.blocks)):
        seqs = S.block_seqs(bi)
        cns, oinfo = orc.generate_consensus(seqs, min_cov, min_idt, want_info=True)
        for j in range(1, len(seqs)):
            g, o = info[p], oinfo[j]
            diffs = []
            for f in FIELDS:
                gv, ov = getattr(g, f), getattr(o, f)
                if f in ("dist", "aln_size", "q_e", "t_e") and not o.aligned:
                    continue
                if gv != ov:
                    diffs.append("%s gpu=%d oracle=%d" % (f, gv, ov))
            if diffs:
                bad_pairs += 1
                if verbose and bad_pairs <= 10:
                    print("  block %d pair %d: %s" % (bi, j, "; ".join(diffs)))
            p += 1
        ok = out[bi] == cns
        if not ok:
            bad_blocks += 1
            if verbose and bad_blocks <= 5:
                n = min(len(out[bi]), len(cns))
                first = next((i for i in range(n) if out[bi][i] != cns[i]), n)
                print("  block %d consensus differs: len gpu=%d oracle=%d first diff at %d" %
                      (bi, len(out[bi]), len(cns), first))
                print("    gpu   ", out[bi][max(0, first - 20):first + 20])
                print("    oracle", cns[max(0, first - 20):first + 20])
    st = eng.stats()
    print("blocks=%d pairs=%d bad_pairs=%d bad_blocks=%d wall=%.3fs  %s" %
          (len(S.blocks), p, bad_pairs, bad_blocks, dt,
           " ".join("%s=%.2f" % (k, v) for k, v in st.items() if k.startswith("ms_"))))
    return bad_pairs == 0 and bad_blocks == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=60000)
    ap.add_argument("--len", type=int, default=5000)
    ap.add_argument("--cov", type=float, default=30)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    eng = Engine(0)
    orc = Oracle()
    ok = True
    S = synth.make_set(a.genome, a.len, a.cov, seed=a.seed, n_blocks=a.blocks)
    ok &= compare(eng, orc, S)
    S = synth.make_set(40000, 3000, 20, seed=a.seed + 1, n_blocks=4, len_sigma=0.4)
    ok &= compare(eng, orc, S, min_cov=2)
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
