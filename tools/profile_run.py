"""Light-weight driver for profiling (no torch import): one synthetic wave through the engine.

    FCX_PROFILE=1 python tools/profile_run.py --blocks 2048 --reps 2
    ncu --set full -k regex:k_dp -c 1 -o gpurun_out/prof_dp python tools/profile_run.py --blocks 512
"""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from falcon_b200 import synth  # noqa: E402
from falcon_b200.binding import Engine, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--read-len", type=int, default=15000)
    ap.add_argument("--cov", type=float, default=50)
    ap.add_argument("--blocks", type=int, default=512)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--max-n-read", type=int, default=200)
    a = ap.parse_args()
    n_reads = int(round(a.genome * a.cov / a.read_len))
    t0 = time.time()
    S = synth.make_set(a.genome, a.read_len, a.cov, n_blocks=a.blocks, max_n_read=a.max_n_read,
                       block_stride=max(1, n_reads // a.blocks))
    print("workload: %d blocks, %d pairs (%.1fs)" % (len(S.blocks), S.n_pairs, time.time() - t0), flush=True)
    eng = Engine(0)
    eng.set_option("pair_info", 0)
    eng.upload_pool(S.pool)
    import numpy as np
    boff = np.zeros(len(S.blocks) + 1, dtype=np.uint32)
    np.cumsum([len(b) for b in S.blocks], out=boff[1:])
    bids = np.concatenate(S.blocks).astype(np.uint32)
    L = lib()
    for r in range(a.reps):
        t0 = time.time()
        eng.consensus_blocks_raw(boff, bids, 4, 0.70, copy=False)
        dt = time.time() - t0
        st = eng.stats()
        print("rep %d wall %.1f ms  pairs/s %.0f  " % (r, dt * 1e3, st["pairs"] / dt) +
              " ".join("%s=%.2f" % (k[3:], v) for k, v in st.items() if k.startswith("ms_")), flush=True)
    print("counters:", {k: v for k, v in st.items() if not k.startswith("ms_")})


if __name__ == "__main__":
    main()
