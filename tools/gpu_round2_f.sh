#!/bin/bash
# GPU session F (round 2): parity suite, default bench line + reference arm, read-length sweep (BASELINE config 4).
set -u
OUT=gpurun_out
TAG=${TAG:-r2f}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -n 3 $OUT/${TAG}_pytest.log
FCX_LANES=1 timeout 300 python tools/profile_run.py --blocks 2960 --reps 3 > $OUT/${TAG}_run.log 2>&1; grep "^rep" $OUT/${TAG}_run.log
timeout 300 python tools/profile_run.py --blocks 8880 --reps 3 > $OUT/${TAG}_run3.log 2>&1; grep "^rep" $OUT/${TAG}_run3.log
timeout 1200 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench_n1.json; tail -n 5 $OUT/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $OUT/${TAG}_bench_ref.json
for rl in 5000 30000 60000; do
  timeout 900 python bench.py --read-len $rl > $OUT/${TAG}_bench_len$rl.json 2> $OUT/${TAG}_bench_len$rl.err; echo "len $rl rc=$?"; cut -c1-300 $OUT/${TAG}_bench_len$rl.json; tail -n 3 $OUT/${TAG}_bench_len$rl.err
done
