#!/bin/bash
# A/B of environment-selected kernel variants on the decode48 line: r2_ab.sh TAG "ENV1=.. ENV2=.." "ENV1=.." ...
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for envs in "$@"; do
  env $envs python bench.py --no-cpu-baseline --no-secondary ${BENCH_ARGS:-} > $OUT/${TAG}_$i.json 2> $OUT/${TAG}_$i.err
  python - <<PY
import json
d = json.load(open("$OUT/${TAG}_$i.json"))
print("$envs", "|", d["config"].get("name"), round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms", {k.split("::")[-1]: round(v, 4) for k, v in d["roofline"].get("kernels_ms", {}).items()})
PY
  i=$((i+1))
done
