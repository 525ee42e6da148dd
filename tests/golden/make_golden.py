"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only where /root/reference exists (the authoring container).  It drives the reference's own
CLI, falcon_kit/mains/consensus.py, under Python 3 through the two-piece shim of SURVEY.md 8(c):
  (1) a fake ``ext_falcon`` module whose __file__ is oracle/_ref/falcon.so (the reference C
      sources compiled by oracle/Makefile), and
  (2) bytes<->str wrappers around get_consensus_without_trim / _with_trim (Py3 c_char_p).
Inputs:  LA4Falcon-format streams built from the reference's test_data/t1.fa, t2.fa and from the
seeded synthetic generator.  Outputs: *.in (stdin), *.out (stdout of the reference CLI).

    python tests/golden/make_golden.py
"""
import hashlib
import io
import json
import os
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

DRIVER = r'''
import sys, types
ref_so, ref_root = sys.argv[1], sys.argv[2]
m = types.ModuleType("ext_falcon"); m.__file__ = ref_so; sys.modules["ext_falcon"] = m
sys.path.insert(0, ref_root)
import falcon_kit.mains.consensus as c
def wrap(f):
    def g(c_input):
        seqs, seed_id, config = c_input
        cns, sid = f(([s.encode() for s in seqs], seed_id, config))
        return cns.decode(), sid
    return g
c.get_consensus_without_trim = wrap(c.get_consensus_without_trim)
c.get_consensus_with_trim = wrap(c.get_consensus_with_trim)
c.main(["consensus"] + sys.argv[3:])
'''


def fasta_seq(path):
    return "".join(l.strip() for l in open(path) if not l.startswith(">"))


def run_ref(stdin_bytes, args):
    so = os.path.join(ROOT, "oracle", "_ref", "falcon.so")
    p = subprocess.run([sys.executable, "-c", DRIVER, so, REF] + args, input=stdin_bytes,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return p.stdout


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    sys.path.insert(0, ROOT)
    from falcon_b200 import synth
    t1 = fasta_seq(os.path.join(REF, "test_data", "t1.fa"))
    t2 = fasta_seq(os.path.join(REF, "test_data", "t2.fa"))
    cases = {}
    block_a = ("00000001 %s\n00000002 %s\n+ +\n- -\n" % (t1, t2)).encode()
    block_b = ("00000002 %s\n00000001 %s\n+ +\n- -\n" % (t2, t1)).encode()
    base = ["--n-core", "0", "--min-n-read", "1", "--min-cov-aln", "0"]
    cases["t1t2_a_default_cov0"] = (block_a, base + ["--min-cov", "0"])
    cases["t1t2_a_default_cov1"] = (block_a, base + ["--min-cov", "1"])
    cases["t1t2_a_multi_cov0"] = (block_a, base + ["--output-multi", "--min-cov", "0"])
    cases["t1t2_a_full_cov1"] = (block_a, base + ["--output-full", "--min-cov", "1"])
    cases["t1t2_b_full_cov1"] = (block_b, base + ["--output-full", "--min-cov", "1"])
    S = synth.make_set(30000, 3000, 25, seed=7, n_blocks=8)
    txt = S.la4falcon_text()
    cases["synth8_multi"] = (txt, ["--n-core", "0", "--output-multi", "--min-idt", "0.70", "--min-cov", "4",
                                   "--max-n-read", "200"])
    cases["synth8_default"] = (txt, ["--n-core", "0", "--min-cov", "4", "--min-n-read", "5"])
    cases["synth8_full_maxn12"] = (txt, ["--n-core", "0", "--output-full", "--min-cov", "2", "--max-n-read", "12",
                                         "--min-cov-aln", "2"])
    cases["synth8_trim_multi"] = (txt, ["--n-core", "0", "--trim", "--output-multi", "--min-cov", "3", "--trim-size", "30",
                                         "--edge-tolerance", "800"])
    S2 = synth.make_set(25000, 2500, 20, seed=11, n_blocks=6, len_sigma=0.5)
    txt2 = S2.la4falcon_text() .replace(b"- -\n", b"junk line with three tokens\n* *\n- -\n")
    cases["synth6_ragged_multi"] = (txt2, ["--n-core", "0", "--output-multi", "--min-cov", "3", "--min-len-aln", "1500",
                                           "--max-cov-aln", "12"])
    manifest = {}
    inputs = {}
    for name, (inp, args) in cases.items():
        out = run_ref(inp, args)
        md5 = hashlib.md5(inp).hexdigest()
        if md5 not in inputs:       # several cases share one stdin stream: store it once
            inputs[md5] = "input_%s.in" % name.rsplit("_", 1)[0] if name.startswith("synth") else "input_%s.in" % name[:6]
            open(os.path.join(HERE, inputs[md5]), "wb").write(inp)
        open(os.path.join(HERE, name + ".out"), "wb").write(out)
        manifest[name] = dict(args=args, input=inputs[md5], in_md5=md5, out_md5=hashlib.md5(out).hexdigest(),
                              out_bytes=len(out), out_lines=out.count(b"\n"))
        print(name, manifest[name]["out_md5"], len(out))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
