"""CPU: the N>1 path -- block partition, per-rank processing and in-order merge -- exercised with
world_size 2 over gloo.  The per-rank engine is the oracle stand-in (no GPU here); the sharding /
merge code is the product's (falcon_b200/shard.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_tiles_and_balances():
    from falcon_b200 import shard
    costs = np.arange(1, 101, dtype=float)
    for w in (1, 2, 3, 8):
        parts = shard.partition(costs, w)
        assert parts[0][0] == 0 and parts[-1][1] == 100
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sums = [costs[a:b].sum() for a, b in parts]
        assert max(sums) < 1.35 * costs.sum() / w + costs.max()
    assert shard.partition([], 4) == [(0, 0)] * 4
    assert shard.merge_in_order([(2, [b"c"]), (0, [b"a", b"b"])]) == [b"a", b"b", b"c"]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from falcon_b200 import shard, synth
    from helpers import OracleEngine
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = synth.make_set(30000, 2500, 15, seed=33, n_blocks=6)
    costs = shard.block_costs([len(b) for b in S.blocks], [len(S.pool[b[0]]) for b in S.blocks])
    lo, hi = shard.partition(costs, world)[rank]
    eng = OracleEngine(Oracle())
    eng.upload_pool(S.pool)
    local = eng.consensus_blocks([S.blocks[i].tolist() for i in range(lo, hi)], 4, 0.70)
    merged = shard.gather_results(local, lo, world)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_merge_to_single_rank_result():
    import multiprocessing as mp
    from falcon_b200 import synth
    from helpers import OracleEngine
    from oracle.oracle import Oracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S = synth.make_set(30000, 2500, 15, seed=33, n_blocks=6)
    eng = OracleEngine(Oracle())
    eng.upload_pool(S.pool)
    assert merged == eng.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)


def _emu_worker(rank, world, port, q):
    """One rank of bench.py's N > 1 data path on the SIMT-emulator build of the engine: its part of the reads
    uploaded and packed (fcx_pool_reserve / upload_part), the other parts received by a broadcast of the packed
    words (gloo here, NCCL on the GPUs), its cost-balanced slice of the seed blocks, ordered gather."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import ctypes as C
    import torch
    import torch.distributed as dist
    from falcon_b200 import shard, synth
    from helpers import emu_engine
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = synth.make_set(30000, 2500, 15, seed=35, n_blocks=6)
    n = len(S.pool)
    cuts = [n * r // world for r in range(world + 1)]
    lens = np.array([len(x) for x in S.pool], dtype=np.uint64)
    all_off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(lens, out=all_off[1:])
    eng = emu_engine()
    eng.pool_reserve(all_off)
    r0, r1 = cuts[rank], cuts[rank + 1]
    blob = np.frombuffer(b"".join(S.pool[r0:r1]), dtype=np.uint8).copy()
    poff = (all_off[r0:r1 + 1] - all_off[r0]).astype(np.uint64)
    eng.pool_upload_part(blob.ctypes.data, poff, r0)
    ptr, n_words, woff = eng.pool_device()
    words = np.ctypeslib.as_array((C.c_int32 * n_words).from_address(ptr))        # emulator: device memory is host memory
    view = torch.from_numpy(words)
    for k in range(world):
        w0, w1 = int(woff[cuts[k]]), int(woff[cuts[k + 1]])
        if w1 > w0:
            dist.broadcast(view[w0:w1], src=k)
    eng.pool_commit()
    costs = shard.block_costs([len(b) for b in S.blocks], [len(S.pool[b[0]]) for b in S.blocks])
    lo, hi = shard.partition(costs, world)[rank]
    local = eng.consensus_blocks([S.blocks[i].tolist() for i in range(lo, hi)], 4, 0.70)
    merged = shard.gather_results(local, lo, world)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pool_parts_broadcast_and_sharded_blocks_emu():
    import multiprocessing as mp
    from falcon_b200 import synth
    from helpers import emu_engine
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    emu_engine()                                   # build the emulator library once, before the ranks start
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_emu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S = synth.make_set(30000, 2500, 15, seed=35, n_blocks=6)
    e = emu_engine()
    e.upload_pool(S.pool)
    assert merged == e.consensus_blocks([b.tolist() for b in S.blocks], 4, 0.70)
