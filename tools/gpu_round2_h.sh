#!/bin/bash
# GPU session H: per-kernel times at 15 kb and 60 kb (single lane), launch list at 60 kb.
set -u
OUT=gpurun_out
TAG=${TAG:-r2h}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -n 2 $OUT/${TAG}_pytest.log
FCX_LANES=1 timeout 300 python tools/profile_run.py --blocks 2960 --reps 3 > $OUT/${TAG}_run15.log 2>&1; grep "^rep" $OUT/${TAG}_run15.log
FCX_LANES=1 FCX_TRACE_WAVES=1 timeout 300 python tools/profile_run.py --blocks 1024 --read-len 60000 --reps 2 > $OUT/${TAG}_run60.log 2>&1; grep "^rep\|^wave" $OUT/${TAG}_run60.log
FCX_LANES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches60.csv python tools/profile_run.py --blocks 512 --read-len 60000 --reps 1 > $OUT/${TAG}_launches60.log 2>&1
python - <<'PY'
import csv,os
rows=[r for r in csv.reader(open(os.path.join("gpurun_out", os.environ.get("TAG","r2h")+"_launches60.csv"))) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4].split("(")[0][-24:], r[7], r[8], r[-2], r[-1])
PY
