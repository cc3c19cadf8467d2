#!/bin/bash
# A/B of library variants on the bench workload (device-timed value + stage times): tools/ab_variants.sh NAME [NAME ...]
# (tools/bin/libgsm_NAME.so from tools/build_variant.sh; the in-tree library runs last)
set -u
for name in "$@" in-tree; do
  if [ "$name" != "in-tree" ]; then export GSM_B200_LIB=$PWD/tools/bin/libgsm_$name.so; else unset GSM_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 --no-modes --no-cpu-baseline > gpurun_out/bench_ab.json 2>/dev/null
  python - "$name" <<'PY'
import json, sys
d = json.loads([l for l in open("gpurun_out/bench_ab.json") if l.startswith("{")][-1])
print(sys.argv[1], round(d["value"], 1), "fps", round(d["ms_per_step"], 4), "ms;", {k: round(v * 1e3, 1) for k, v in d["stage_ms"].items()})
PY
done
