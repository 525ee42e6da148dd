"""Randomised parity sweep of the kernel sources on the SIMT emulator (tests/emu) against the CPU oracle:
random genome / read length / coverage / error mix / min_cov / min_idt, stage outputs and consensus compared.
TEST INFRASTRUCTURE (no GPU needed).   python tools/emu_stress.py [n_cases] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from falcon_b200 import synth  # noqa: E402
from helpers import emu_engine  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
import test_gpu_parity as G  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    e, orc = emu_engine(), Oracle()
    for c in range(n):
        rl = int(rng.choice([1200, 2500, 4000, 7000]))
        p = dict(genome_size=int(rl * rng.integers(6, 14)), read_len=rl, coverage=float(rng.choice([8, 15, 25, 40])),
                 seed=int(rng.integers(1, 1 << 30)), n_blocks=int(rng.integers(1, 4)),
                 p_ins=float(rng.choice([0.02, 0.06, 0.09, 0.13])), p_del=float(rng.choice([0.01, 0.045, 0.09])),
                 p_sub=float(rng.choice([0.0, 0.015, 0.04])), len_sigma=float(rng.choice([0.0, 0.0, 0.35])),
                 max_n_read=int(rng.choice([12, 60, 200])))
        min_cov, min_idt = int(rng.choice([0, 1, 4, 8])), float(rng.choice([0.6, 0.7, 0.8]))
        S = synth.make_set(**p)
        G._check_set(e, orc, S, min_cov, min_idt)
        print("case %d ok: %s min_cov %d min_idt %.2f (%d pairs)" % (c, p, min_cov, min_idt, S.n_pairs), flush=True)
    print("ALL %d CASES OK" % n)


if __name__ == "__main__":
    main()
