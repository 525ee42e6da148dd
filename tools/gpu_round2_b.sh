#!/bin/bash
# GPU session B (round 2): parity suite, per-kernel times, default bench line (N=1) + reference arm,
# ncu captures of the new consensus kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r2b_pytest.log
tail -n 3 $OUT/r2b_pytest.log
FCX_LANES=1 timeout 300 python tools/profile_run.py --blocks 2960 --reps 2 > $OUT/r2b_run.log 2>&1; tail -n 4 $OUT/r2b_run.log
timeout 1200 python bench.py > $OUT/r2b_bench_n1.json 2> $OUT/r2b_bench_n1.err; echo "bench rc=$?"; tail -c 3000 $OUT/r2b_bench_n1.json; tail -n 5 $OUT/r2b_bench_n1.err
timeout 600 python bench.py --impl reference > $OUT/r2b_bench_ref.json 2> $OUT/r2b_bench_ref.err; echo "ref rc=$?"; tail -c 1500 $OUT/r2b_bench_ref.json
export FCX_LANES=1
for k in k_vote k_cns_dp k_traceback; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/r2b_full_$k python tools/profile_run.py --blocks 2960 --reps 1 > $OUT/r2b_full_$k.log 2>&1
done
ls -la $OUT | tail -n 12
