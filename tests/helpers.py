"""Shared test helpers: an oracle-backed stand-in engine for CPU tests of the host logic."""
import io
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleEngine:
    """Same surface as falcon_b200.binding.Engine, arithmetic by the CPU oracle (tests only)."""

    def __init__(self, oracle, ref=None):
        self.oracle = oracle
        self.ref = ref          # compiled reference: needed only for --trim (k-mer chaining)
        self.pool = []

    def upload_pool(self, reads):
        self.pool = list(reads)

    def consensus_blocks(self, blocks, min_cov, min_idt, K=8):
        return [self.oracle.generate_consensus([self.pool[i] for i in b], min_cov, min_idt, K) for b in blocks]

    # raw (pointer / numpy) surface used by the native-parser path of the CLI
    def upload_pool_raw(self, bases_ptr, offsets):
        import ctypes as C
        self.pool = [C.string_at(bases_ptr + int(offsets[i]), int(offsets[i + 1] - offsets[i]))
                     for i in range(len(offsets) - 1)]

    def consensus_blocks_raw(self, block_off, read_ids, min_cov, min_idt, K=8):
        import numpy as np
        blocks = [read_ids[int(block_off[b]):int(block_off[b + 1])].tolist() for b in range(len(block_off) - 1)]
        cns = self.consensus_blocks(blocks, min_cov, min_idt, K)
        off = np.zeros(len(cns) + 1, dtype=np.uint64)
        np.cumsum([len(c) for c in cns], out=off[1:])
        return np.frombuffer(b"".join(cns), dtype=np.uint8), off


    def trim_blocks_raw(self, block_off, read_ids, edge_tolerance, trim_size, max_n_read, max_cov_aln):
        """Stand-in for fcx_trim_blocks: the reference's own logic (tests/ref_host.py)."""
        import numpy as np
        import ref_host
        cfg = (0, 8, max_n_read, 0.7, edge_tolerance, trim_size, 0, max_cov_aln)
        new_off, new_ids = [0], []
        for b in range(len(block_off) - 1):
            seqs = [self.pool[i] for i in read_ids[int(block_off[b]):int(block_off[b + 1])]]
            out = ref_host.trim_block(self.ref, seqs, cfg)
            new_ids.append(int(read_ids[int(block_off[b])]))
            for sq in out[1:]:
                self.pool.append(sq)
                new_ids.append(len(self.pool) - 1)
            new_off.append(len(new_ids))
        return np.asarray(new_off, dtype=np.uint32), np.asarray(new_ids, dtype=np.uint32)


def golden_cases():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def run_cli(argv, stdin_bytes, engine):
    from falcon_b200 import consensus
    args = consensus.parse_args(["consensus"] + list(argv))
    out = io.StringIO()
    consensus.run(args, stdin=io.BytesIO(stdin_bytes), stdout=out, engine=engine)
    return out.getvalue().encode()


def run_cli_python_host(argv, stdin_bytes, engine, ref=None):
    """The same CLI contract with the host side done by tests/ref_host.py (Python restatement of the
    reference's parser / read selection / trim): the checker for the native parser path."""
    import ref_host
    from falcon_b200 import consensus
    args = consensus.parse_args(["consensus"] + list(argv))
    cfg = (args.min_cov, 8, args.max_n_read, args.min_idt, args.edge_tolerance, args.trim_size,
           args.min_cov_aln, args.max_cov_aln)
    out = io.StringIO()
    for seqs, seed_id in ref_host.get_seq_data(io.BytesIO(stdin_bytes), cfg, args.min_n_read, args.min_len_aln):
        if args.trim:
            seqs = ref_host.trim_block(ref, seqs, cfg)
        elif len(seqs) > args.max_n_read:
            seqs = ref_host.get_longest_reads(seqs, args.max_n_read, args.max_cov_aln, sort=True)
        engine.upload_pool(seqs)
        cns = engine.consensus_blocks([list(range(len(seqs)))], args.min_cov, args.min_idt)[0]
        consensus.emit(out, cns.decode(), seed_id, args)
    return out.getvalue().encode()


# ---------------------------------------------------------------------------------------------
# SIMT-emulator build of the CUDA sources (tests/emu): lets the kernel logic be checked against the
# oracle on a machine without a GPU.  Test infrastructure only -- falcon_b200.binding never loads it.
EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libfalcon_b200_emu.so")


def build_emu():
    import subprocess
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    return EMU_SO


def emu_engine():
    """A falcon_b200.binding.Engine whose library handle is the emulator build."""
    import ctypes as C
    from falcon_b200 import binding

    class EmuEngine(binding.Engine):
        def __init__(self):
            self._lib = binding.load_library(build_emu())
            h = C.c_void_p()
            if self._lib.fcx_create(0, C.byref(h)) != 0:
                raise binding.EngineError("fcx_create (emu): %s" % self._lib.fcx_last_error(None).decode())
            self._h = h
            self.n_reads = 0

    return EmuEngine()


def emu_multi_engine(n_devices=2):
    """falcon_b200.binding.MultiEngine on the emulator build (FCX_EMU_DEVICES fake devices)."""
    from falcon_b200 import binding
    os.environ["FCX_EMU_DEVICES"] = str(n_devices)

    class EmuMulti(binding.MultiEngine):
        def __init__(self, devices):
            self._lib = binding.load_library(build_emu())
            self._open(devices)

    return EmuMulti(list(range(n_devices)))
