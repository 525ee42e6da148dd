"""The C ABI from plain C: tests/c/abi_smoke.c is compiled against include/falcon_b200.h and linked
with libfalcon_b200.so (CPU: it must compile and link; GPU: it must give the oracle's consensus)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "falcon_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", LIBDIR, "-lfalcon_b200", "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def test_c_caller_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run(["nm", "-u", exe], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "generate_consensus" in out and "free_consensus_data" in out


@pytest.mark.gpu
def test_c_caller_matches_oracle(tmp_path, oracle):
    from falcon_b200 import synth
    exe = _build(tmp_path)
    S = synth.make_set(30000, 3000, 20, seed=51, n_blocks=1)
    seqs = S.block_seqs(0)
    p = tmp_path / "seqs.txt"
    p.write_bytes(b"\n".join(seqs) + b"\n")
    r = subprocess.run([exe, str(p), "4", "0.70"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    want, eqv = oracle.generate_consensus(seqs, 4, 0.70, want_eqv=True)
    assert r.stdout.strip() == want
    assert ("eqv_sum=%d" % sum(eqv)).encode() in r.stderr
