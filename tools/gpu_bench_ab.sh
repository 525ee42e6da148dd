#!/bin/bash
# several bench.py runs in one session (same box): run-to-run spread and lane-count variants
set -u
OUT=gpurun_out
TAG=${TAG:-r2l}
mkdir -p $OUT
i=0
for env in "FCX_LANES=3" "FCX_LANES=3" "FCX_LANES=2" "FCX_LANES=4" "FCX_LANES=3 FCX_WAVE_BLOCKS=2220"; do
  i=$((i+1))
  env $env timeout 600 python bench.py --no-cpu-baseline --no-e2e > $OUT/${TAG}_ab$i.json 2> $OUT/${TAG}_ab$i.err
  python - "$env" $OUT/${TAG}_ab$i.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2]))
print(sys.argv[1], "value %.0f ms/step %.1f"%(d["value"], d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
PY
done
