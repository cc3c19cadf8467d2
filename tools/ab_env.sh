#!/bin/bash
# A/B by environment on a bench workload: tools/ab_env.sh [-w WORKLOAD] VAR=VAL ...   (one run per setting, then the default)
W=C2
if [ "${1:-}" == "-w" ]; then W=$2; shift 2; fi
for kv in "$@" ""; do
  name=${kv:-default}
  if [ -n "$kv" ]; then export "$kv"; fi
  python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --no-modes > gpurun_out/bench_ab.json 2>/dev/null
  python - "$name" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 1), "fps", round(d["ms_per_step"], 4), "ms;", {k: round(v * 1e3, 1) for k, v in d["stage_ms"].items()})
PY
  if [ -n "$kv" ]; then unset "${kv%%=*}"; fi
done
