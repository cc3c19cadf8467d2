#!/bin/bash
# A/B of builds of the library on the bench workload: tools/ab_bench.sh <a.so> [<b.so> ...]   (the in-tree library runs last)
set -u
for lib in "$@" ""; do
  name=${lib:-in-tree}
  if [ -n "$lib" ]; then export GSM_B200_LIB=$lib; else unset GSM_B200_LIB; fi
  python bench.py --steps 20 --warmup 5 > gpurun_out/bench_ab.json
  python - "$name" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/bench_ab.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"], 1), "fps", round(d["ms_per_step"], 4), "ms;", {k: round(v * 1e3, 1) for k, v in d["stage_ms"].items()}, "e2e", round(d["e2e"]["value"], 1))
PY
done
