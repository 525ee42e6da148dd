#!/bin/bash
# GPU session A (round 2): parity suite, A/B of the DP kernels, ncu capture of k_dp3.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r2a_pytest.log
tail -5 $OUT/r2a_pytest.log
export FCX_LANES=1
FCX_DP_VARIANT=3 timeout 300 python tools/profile_run.py --blocks 2960 --reps 2 > $OUT/r2a_run_v3.log 2>&1
FCX_DP_VARIANT=1 timeout 300 python tools/profile_run.py --blocks 2960 --reps 2 > $OUT/r2a_run_v1.log 2>&1
tail -4 $OUT/r2a_run_v3.log $OUT/r2a_run_v1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dp3 -c 1 -f -o $OUT/r2a_full_k_dp3 python tools/profile_run.py --blocks 2960 --reps 1 > $OUT/r2a_full_k_dp3.log 2>&1
ls -la $OUT | tail -8
