#!/bin/bash
# GPU session C (round 2): parity suite, per-kernel times, ncu captures of the rewritten kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG:-r2c}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG:-r2c}_pytest.log
tail -n 3 $OUT/${TAG:-r2c}_pytest.log
FCX_LANES=1 timeout 300 python tools/profile_run.py --blocks 2960 --reps 3 > $OUT/${TAG:-r2c}_run.log 2>&1; tail -n 5 $OUT/${TAG:-r2c}_run.log
timeout 300 python tools/profile_run.py --blocks 8880 --reps 3 > $OUT/${TAG:-r2c}_run2.log 2>&1; tail -n 5 $OUT/${TAG:-r2c}_run2.log
export FCX_LANES=1
for k in ${PROFILE_KERNELS:-k_dp3 k_cns_dp k_traceback k_range}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/${TAG:-r2c}_full_$k python tools/profile_run.py --blocks 2960 --reps 1 > $OUT/${TAG:-r2c}_full_$k.log 2>&1
done
ls -la $OUT | tail -n 12
