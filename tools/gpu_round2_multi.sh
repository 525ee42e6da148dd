#!/bin/bash
# Multi-GPU session (round 2): N = $1 GPUs of one box.  Parity suite (fcx_multi on all GPUs), the bench line at N
# (one data set sharded over the N ranks, strong scaling), optionally the D. mel-sized run (BASELINE config 3).
set -u
N=${1:-2}
OUT=gpurun_out
TAG=${TAG:-r2m}
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_n${N}_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "multi or stage_parity" > $OUT/${TAG}_n${N}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_n${N}_pytest.log
tail -n 3 $OUT/${TAG}_n${N}_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench N=$N rc=$?"; cut -c1-600 $OUT/${TAG}_bench_n$N.json; tail -n 4 $OUT/${TAG}_bench_n$N.err
if [ "${DMEL:-0}" = "1" ]; then
  timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 2 --warmup 3 --genome 140000000 --no-e2e > $OUT/${TAG}_bench_dmel_n$N.json 2> $OUT/${TAG}_bench_dmel_n$N.err
  echo "dmel N=$N rc=$?"; cut -c1-600 $OUT/${TAG}_bench_dmel_n$N.json; tail -n 4 $OUT/${TAG}_bench_dmel_n$N.err
fi
if [ "${ONEPROC:-0}" = "1" ]; then
  # one process owning all N GPUs (fcx_multi): the drop-in CLI's --devices path
  timeout 600 python tools/bench_text.py --blocks 8192 --streams 4 --devices 0-$((N-1)) > $OUT/${TAG}_text_n$N.json 2> $OUT/${TAG}_text_n$N.err
  echo "text N=$N rc=$?"; cat $OUT/${TAG}_text_n$N.json; tail -n 3 $OUT/${TAG}_text_n$N.err
fi
