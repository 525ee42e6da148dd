#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one bench-like step + one full ncu capture per
# hot kernel.  Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/profile_run.py --blocks ${PROFILE_BLOCKS:-2960} --reps 1"
export FCX_LANES=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv $CMD > $OUT/launches.log 2>&1
for k in ${PROFILE_KERNELS:-k_dp k_consensus k_traceback k_range k_transpose k_index}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $OUT/full_$k $CMD > $OUT/full_$k.log 2>&1
done
ls -la $OUT
