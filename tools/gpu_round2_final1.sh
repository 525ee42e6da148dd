#!/bin/bash
# Final single-GPU session of round 2: parity suite, launch list, one `ncu --set full` capture per hot kernel
# (+ the TMA-staged DP variant), compute-sanitizer, bench line + reference arm, read-length sweep, text path.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/final_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/final_pytest.log
tail -n 2 $OUT/final_pytest.log
CMD="python tools/profile_run.py --blocks 2960 --reps 1"
export FCX_LANES=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
for k in k_dp3 k_vote k_cns_dp k_traceback k_range k_index; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/full_$k $CMD > $OUT/full_$k.log 2>&1
done
FCX_DP_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dp -c 1 -f -o $OUT/full_k_dp_staged $CMD > $OUT/full_k_dp_staged.log 2>&1
FCX_DP_VARIANT=2 timeout 300 $CMD --reps 3 > $OUT/final_run_staged.log 2>&1; grep "^rep" $OUT/final_run_staged.log | tail -1
FCX_DP_VARIANT=1 timeout 300 $CMD --reps 3 > $OUT/final_run_kdp.log 2>&1; grep "^rep" $OUT/final_run_kdp.log | tail -1
timeout 300 $CMD --reps 3 > $OUT/final_run.log 2>&1; grep "^rep" $OUT/final_run.log | tail -1
unset FCX_LANES
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/sanitize_racecheck.log
timeout 1200 python bench.py > $OUT/final_bench_n1.json 2> $OUT/final_bench_n1.err; echo "bench rc=$?"; cut -c1-200 $OUT/final_bench_n1.json
timeout 600 python bench.py --impl reference > $OUT/final_bench_ref.json 2> $OUT/final_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/final_bench_ref.json
for rl in 5000 30000 60000; do
  timeout 900 python bench.py --read-len $rl > $OUT/final_bench_len$rl.json 2> $OUT/final_bench_len$rl.err; echo "len $rl rc=$?"; cut -c1-200 $OUT/final_bench_len$rl.json
done
timeout 900 python tools/bench_text.py --blocks 4096 --streams 1,4,8 > $OUT/final_text_n1.json 2> $OUT/final_text_n1.err; echo "text rc=$?"; cat $OUT/final_text_n1.json
