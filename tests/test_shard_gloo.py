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
