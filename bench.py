#!/usr/bin/env python
"""bench.py -- aligned read-pairs/s of the fc_consensus hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the whole hot path (k-mer range -> banded O(ND) DP -> traceback ->
alignment-graph consensus) over this rank's seed blocks of the synthetic workload:
E. coli-like uniform random genome, 50x coverage of 15 kb reads, 15 % error (ins 9 / del 4.5 /
sub 1.5), blocks built from ground truth.  Weak scaling: every rank owns an equal
slice of a genome that grows with N; there is no data-path collective.

`value`  : pairs/s with the read pool already resident in HBM (timed with CUDA events on the
           engine's stream, max over ranks).
`e2e`    : the same metric through the public batched C-ABI call with HOST buffers: every step
           uploads the read pool from pinned host memory and reads the consensus back.
`roofline`: dominant kernel, algorithmic bytes / its CUDA-event time.
`cpu_baseline`: the reference's own C code (oracle/_ref/falcon.so, or the oracle port if that is
           absent) on the host cores, on a bounded sample of the same blocks.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "aligned read-pairs/sec fc_consensus"
UNIT = "pairs/s"


# ------------------------------------------------------------------------------- workload
def build_workload(args, rank):
    from falcon_b200 import synth
    n_reads = int(round(args.genome * args.cov / args.read_len))
    stride = max(1, n_reads // args.blocks) if args.blocks and args.blocks < n_reads else 1
    S = synth.make_set(args.genome, args.read_len, args.cov, seed=args.seed + 1000 * rank,
                       n_blocks=args.blocks if args.blocks else None, max_n_read=args.max_n_read,
                       block_stride=stride)
    return S


def flatten(S, pinned=True):
    """Pool -> (pinned uint8 buffer, uint64 offsets), blocks -> (uint32 block_off, uint32 read_ids)."""
    from falcon_b200.binding import PinnedBuffer
    lens = np.fromiter((len(r) for r in S.pool), dtype=np.uint64, count=len(S.pool))
    off = np.zeros(len(S.pool) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    buf = PinnedBuffer(total)
    cat = np.frombuffer(b"".join(S.pool), dtype=np.uint8)
    buf.array[:total] = cat
    block_off = np.zeros(len(S.blocks) + 1, dtype=np.uint32)
    np.cumsum([len(b) for b in S.blocks], out=block_off[1:])
    ids = np.concatenate(S.blocks).astype(np.uint32)
    return buf, off, block_off, ids


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU reference
_W = {}


def _ref_worker_init(so_path, kind):
    import ctypes as C
    sys.path.insert(0, ROOT)
    if kind == "reference":
        from oracle.oracle import Ref
        _W["eng"] = Ref(so_path)
    else:
        from oracle.oracle import Oracle
        _W["eng"] = Oracle()


def _ref_worker_run(job):
    seqs, min_cov, min_idt = job
    cns = _W["eng"].generate_consensus(seqs, min_cov, min_idt)
    return len(cns)


def cpu_reference_pool(cores):
    from oracle import oracle as orc
    try:
        orc.build()
    except Exception:
        pass
    kind = "reference" if os.path.exists(orc.REF_SO) else "port"
    ctx = mp.get_context("fork")
    pool = ctx.Pool(cores, initializer=_ref_worker_init, initargs=(orc.REF_SO, kind))
    return pool, kind


def cpu_reference_time(pool, S, block_ids, min_cov, min_idt, chunksize=1):
    jobs = [(S.block_seqs(b), min_cov, min_idt) for b in block_ids]
    pairs = sum(len(j[0]) - 1 for j in jobs)
    t0 = time.perf_counter()
    list(pool.imap(_ref_worker_run, jobs, chunksize))
    dt = time.perf_counter() - t0
    return pairs, dt


# ------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=4_600_000)
    ap.add_argument("--read-len", type=int, default=15000)
    ap.add_argument("--cov", type=float, default=50.0)
    ap.add_argument("--blocks", type=int, default=0,
                    help="seed blocks per rank per step (0 = the whole set: every read is a seed, ~15.3k blocks)")
    ap.add_argument("--max-n-read", type=int, default=200)
    ap.add_argument("--min-cov", type=int, default=4)
    ap.add_argument("--min-idt", type=float, default=0.70)
    ap.add_argument("--seed", type=int, default=20260924)
    ap.add_argument("--cpu-sample-blocks", type=int, default=0, help="blocks in the CPU baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = ("synthetic E. coli-like %.1f Mb, %gx %d kb reads, 15%% error (ins 9/del 4.5/sub 1.5); %s seed blocks "
                "per rank per step, max_n_read %d" % (args.genome / 1e6, args.cov, args.read_len // 1000,
                                                       args.blocks or "all", args.max_n_read))
    config = {"workload": workload, "min_cov": args.min_cov, "min_idt": args.min_idt, "K": 8,
              "parallelism": "seed-block shards, %d rank(s), no data-path collective" % world,
              "l2": "inputs_larger_than_L2"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        S = build_workload(args, 0)
        nblk = args.cpu_sample_blocks or min(len(S.blocks), 2 * cores)
        ids = list(range(nblk))
        pool, kind = cpu_reference_pool(cores)
        for _ in range(max(1, args.warmup)):
            cpu_reference_time(pool, S, ids[:max(1, min(nblk, cores))] if _ == 0 else ids, args.min_cov, args.min_idt)
        t_tot, pairs_tot = 0.0, 0
        for _ in range(args.steps):
            pairs, dt = cpu_reference_time(pool, S, ids, args.min_cov, args.min_idt)
            t_tot += dt; pairs_tot += pairs
        pool.terminate()
        v = pairs_tot / t_tot
        sample = "%d seed blocks (%d pairs) of the workload per step" % (nblk, pairs_tot // max(1, args.steps))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; falcon_b200 has no CPU path", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from falcon_b200.binding import Engine, lib
    import ctypes as C

    S = build_workload(args, rank)
    buf, off, block_off, ids = flatten(S)
    n_pairs = int(S.n_pairs)
    eng = Engine(local_rank)
    eng.set_option("pair_info", 0)
    L = lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn, steps):
        barrier()
        L.fcx_timer_start(eng._h)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        ms = C.c_double()
        L.fcx_timer_stop(eng._h, C.byref(ms))
        wall = time.perf_counter() - t0
        barrier()
        return allmax(ms.value), allmax(wall * 1e3)

    # resident-pool path
    eng.upload_pool_raw(buf.ptr, off)
    step_resident = lambda: eng.consensus_blocks_raw(block_off, ids, args.min_cov, args.min_idt)  # noqa: E731
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    acc = {k: 0.0 for k in ("ms_index", "ms_range", "ms_dp", "ms_traceback", "ms_consensus", "ms_total")}
    cnt = {}

    def step_resident_stats():
        step_resident()
        st = eng.stats()
        for k in acc:
            acc[k] += st[k]
        cnt.update({k: v for k, v in st.items() if not k.startswith("ms_")})

    dev_ms, wall_ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    total_pairs = allsum(float(n_pairs))
    value = total_pairs * args.steps / (dev_ms / 1e3)
    # per-kernel CUDA-event times for the roofline: one extra (untimed) step with a single wave in
    # flight, so that the kernel durations are not inflated by overlapping lanes
    eng.set_option("lanes", 1)
    step_resident_stats()
    eng.set_option("lanes", 0)
    n_kstat = 1

    # e2e path: upload from pinned host memory + consensus + results back, every step
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            eng.upload_pool_raw(buf.ptr, off)
            return eng.consensus_blocks_raw(block_off, ids, args.min_cov, args.min_idt)
        data, ooff = step_e2e()
        for _ in range(max(0, args.warmup - 1)):
            step_e2e()
        e_ms, e_wall = timed(step_e2e, args.steps)
        h2d = int(off[-1]) + off.nbytes * 2 + block_off.nbytes + ids.nbytes
        d2h = int(data.nbytes + ooff.nbytes)
        e2e = {"value": total_pairs * args.steps / (e_wall / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e_wall / args.steps, "api": "fcx_pool_upload + fcx_consensus_blocks"}

    # roofline of the dominant kernel (algorithmic bytes per SURVEY.md 8(d))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kms = {k[3:]: acc[k] / n_kstat for k in acc if k != "ms_total"}
    dom = max(kms, key=kms.get)
    E, D1, A, SP = cnt.get("trace_cells", 0), cnt.get("dp_steps", 0), cnt.get("aln_cols", 0), cnt.get("span_bases", 0)
    seed_bases = float(sum(len(S.pool[b[0]]) for b in S.blocks))
    bytes_dp = SP / 4.0 + 4.0 * E + 8.0 * D1 + 8.0 * A
    bytes_cns = 2 * 8.0 * A + seed_bases
    alg = {"dp": bytes_dp, "consensus": bytes_cns, "traceback": 4.0 * E / 8 + 8.0 * A, "range": SP / 4.0, "index": seed_bases * 4}
    ach = alg.get(dom, 0.0) / (kms[dom] / 1e3) / 1e9 if kms[dom] > 0 else 0.0
    traffic = None
    try:   # dram__bytes_read+write per launch of the dominant kernel, from the committed ncu capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        k = prof.get("k_" + dom)
        if k:
            traffic = {"bytes_per_launch": k["dram_bytes"], "pairs_in_launch": k.get("pairs"),
                       "bytes_per_pair": k["dram_bytes"] / max(1, k.get("pairs", 1)), "source": "profiles/ncu_summary.json"}
    except Exception:
        pass
    kname = {"dp": "k_dp2"}.get(dom, "k_" + dom)      # the default DP kernel is the two-pairs-per-warp k_dp2
    roofline = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic,
                "kernel_timing": "CUDA events on the launching stream, one step with a single wave in flight",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "kernel_ms_per_step": kms, "algorithmic_bytes_per_step": alg[dom],
                "dp_kernel": {"achieved": bytes_dp / (kms["dp"] / 1e3) / 1e9 if kms["dp"] > 0 else 0.0,
                              "frac": (bytes_dp / (kms["dp"] / 1e3) / 1e9 / peak) if kms["dp"] > 0 else 0.0}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pool, kind = cpu_reference_pool(cores)
        nblk = args.cpu_sample_blocks or min(len(S.blocks), 2 * cores)
        cpu_reference_time(pool, S, list(range(min(nblk, cores))), args.min_cov, args.min_idt)   # warm the workers
        pairs, dt = cpu_reference_time(pool, S, list(range(nblk)), args.min_cov, args.min_idt)
        pool.terminate()
        cpu_baseline = {"value": pairs / dt, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": "%d seed blocks (%d pairs) of the same workload, %.1f s wall" % (nblk, pairs, dt)}

    if rank == 0:
        aligned_bases = cnt.get("aln_cols", 0)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "pairs_per_step_per_rank": n_pairs, "wall_ms_per_step": wall_ms / args.steps,
                "gbases_per_s_input": value * args.read_len / 1e9,
                "aligned_columns_per_step_rank0": aligned_bases,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(cnt.get("kernel_launches", 0)) * args.steps,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
