#!/bin/bash
# GPU session E (round 2): host timeline of the waves, lanes 1/2/3, default bench line.
set -u
OUT=gpurun_out
TAG=${TAG:-r2e}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -n 3 $OUT/${TAG}_pytest.log
for lanes in 1 2 3; do
  FCX_TRACE_WAVES=1 FCX_LANES=$lanes timeout 300 python tools/profile_run.py --blocks 8880 --reps 3 > $OUT/${TAG}_lanes$lanes.log 2>&1
  grep "^rep" $OUT/${TAG}_lanes$lanes.log
done
timeout 1200 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -c 2500 $OUT/${TAG}_bench_n1.json; tail -n 5 $OUT/${TAG}_bench_n1.err
