#!/usr/bin/env python
"""the UNMODIFIED reference CLI (falcon_kit/mains/consensus.py through the
Python-3 shim of tests/golden/make_golden.py) timed on LA4Falcon text of the bench workload.

Runs only where /root/reference exists (the authoring container: no GPU, 8 vCPUs); the GPU box has no
reference tree, so bench.py times the reference's C code there (oracle/_ref/falcon.so) and this script
records what the full Python CLI adds on top: start-up (msa_array construction per worker), parsing and
the imap pipe.  Writes one JSON object to stdout.

    python tools/ref_cli_timing.py [--blocks 48] [--cores 1,8]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the shim of tests/golden/make_golden.py with module-level wrappers, so that the reference's
# multiprocessing.Pool (--n-core >= 1) can pickle the worker function
DRIVER = r'''
import sys, types
ref_so, ref_root = sys.argv[1], sys.argv[2]
m = types.ModuleType("ext_falcon"); m.__file__ = ref_so; sys.modules["ext_falcon"] = m
sys.path.insert(0, ref_root)
import falcon_kit.mains.consensus as c
_orig_without, _orig_with = c.get_consensus_without_trim, c.get_consensus_with_trim
def g_without(c_input):
    seqs, seed_id, config = c_input
    cns, sid = _orig_without(([s.encode() for s in seqs], seed_id, config))
    return cns.decode(), sid
def g_with(c_input):
    seqs, seed_id, config = c_input
    cns, sid = _orig_with(([s.encode() for s in seqs], seed_id, config))
    return cns.decode(), sid
c.get_consensus_without_trim = g_without
c.get_consensus_with_trim = g_with
c.main(["consensus"] + sys.argv[3:])
'''
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=48)
    ap.add_argument("--cores", default="1,8")
    a = ap.parse_args()
    import make_golden as G
    from falcon_b200 import synth
    if not os.path.isdir(G.REF):
        print(json.dumps({"unavailable": "no reference tree at %s" % G.REF}))
        return 0
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    S = synth.make_set(4_600_000, 15000, 50, n_blocks=a.blocks, max_n_read=200, block_stride=max(1, 15333 // a.blocks))
    txt = S.la4falcon_text()
    pairs = S.n_pairs
    one = S.la4falcon_text([0])
    opts = ["--output-multi", "--min-idt", "0.70", "--min-cov", "4", "--max-n-read", "200"]
    so = os.path.join(ROOT, "oracle", "_ref", "falcon.so")

    def run(stdin, n_core):
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, "-c", DRIVER, so, G.REF, "--n-core", str(n_core)] + opts, input=stdin,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        return time.perf_counter() - t0, p.stdout

    res = {"what": "unmodified reference CLI (falcon_kit/mains/consensus.py via the Py3 shim) on LA4Falcon text",
           "workload": "synthetic E. coli-like 4.6 Mb, 50x 15 kb, 15%% error: %d seed blocks, %d pairs, %.2f GB of text" %
                       (len(S.blocks), pairs, len(txt) / 1e9),
           "cpu_model": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?"),
           "host_cores": len(os.sched_getaffinity(0)), "runs": []}
    ref_out = None
    for c in [int(x) for x in a.cores.split(",")]:
        startup, _ = run(one, c)                                   # one block: start-up + msa_array per worker
        wall, out = run(txt, c)
        if ref_out is None:
            ref_out = out
        assert out == ref_out, "reference CLI output depends on --n-core?"
        pairs_one = len(S.blocks[0]) - 1
        res["runs"].append({"n_core": c, "wall_s": wall, "pairs_per_s_wall": pairs / wall,
                            "one_block_run_s": startup,
                            "pairs_per_s_steady": (pairs - pairs_one) / max(1e-9, wall - startup)})
    # the same text through our host logic driven by the reference C code must give the same bytes
    import io
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import OracleEngine
    from oracle.oracle import Ref
    from falcon_b200 import consensus
    got = io.StringIO()
    consensus.run(consensus.parse_args(["consensus"] + opts), stdin=io.BytesIO(txt), stdout=got, engine=OracleEngine(Ref()))
    res["host_logic_parity"] = "falcon_b200.consensus host side + reference C == reference CLI: %s" % (got.getvalue().encode() == ref_out)
    import hashlib
    res["fasta_md5"] = hashlib.md5(ref_out).hexdigest()
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
