#!/bin/bash
# Last session of round 2: smoke(), the whole gpu suite, launch list + k_dp3 capture of the final build, bench + reference arm.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/final2_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > $OUT/final_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/final_pytest.log
tail -n 2 $OUT/final_pytest.log
CMD="python tools/profile_run.py --blocks 2960 --reps 1"
export FCX_LANES=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
for k in ${PROFILE_KERNELS:-k_dp3}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/full_$k $CMD > $OUT/full_$k.log 2>&1
done
timeout 300 $CMD --reps 3 > $OUT/final_run.log 2>&1; grep "^rep" $OUT/final_run.log | tail -1
unset FCX_LANES
timeout 1200 python bench.py > $OUT/final_bench_n1.json 2> $OUT/final_bench_n1.err; echo "bench rc=$?"; cut -c1-200 $OUT/final_bench_n1.json
timeout 600 python bench.py --impl reference > $OUT/final_bench_ref.json 2> $OUT/final_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/final_bench_ref.json
