#!/bin/bash
# A/B on one box: wave size and k_cns_dp register cap (rebuilds the library on the box)
set -u
OUT=gpurun_out; TAG=${TAG:-r2o}; mkdir -p $OUT
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-e2e > $OUT/${TAG}_$label.json 2> $OUT/${TAG}_$label.err
  python - "$label" $OUT/${TAG}_$label.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(sys.argv[1], "value %.0f ms/step %.1f"%(d["value"], d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -n 1 $OUT/${TAG}_pytest.log
run base X=1
run wave4440 FCX_WAVE_BLOCKS=4440
run wave5920 FCX_WAVE_BLOCKS=5920
touch falcon_b200/csrc/fcx_engine.cu; make -s -C falcon_b200/csrc EXTRA=-DCDP_MIN_CTAS=6 > /dev/null 2>&1
run cta6 X=1
run cta6_wave3552 FCX_WAVE_BLOCKS=3552
run cta6_wave5920 FCX_WAVE_BLOCKS=5920
