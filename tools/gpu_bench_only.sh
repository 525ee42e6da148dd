#!/bin/bash
set -u
OUT=gpurun_out
TAG=${TAG:-r2k}
mkdir -p $OUT
timeout 1200 python bench.py ${BENCH_ARGS:-} > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-330 $OUT/${TAG}_bench_n1.json; tail -n 5 $OUT/${TAG}_bench_n1.err
python - <<'PY'
import json,os
d=json.load(open(os.path.join("gpurun_out", os.environ.get("TAG","r2k")+"_bench_n1.json")))
print({k:round(v,1) for k,v in d["roofline"]["kernel_ms_per_step"].items()}, "e2e %.0f"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["dp_kernel"]["frac"])
PY
