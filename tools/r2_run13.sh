#!/bin/bash
# validation + timing of a kernel change: all GPU tests, decode48 and roundtrip48 bench lines
set -u
TAG=${1:-r2_v13}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest.log
python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_decode48.json 2> $OUT/${TAG}_bench_decode48.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_decode48.json"))
print("decode48", round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]/1e6, 2), d["roofline"]["kernels_ms"])
for k, v in d.get("secondary", {}).items(): print(" ", k, round(v["ms_per_step"], 4), "ms", round(v["value"]/1e6, 2), "M")
PY
python bench.py --workload roundtrip48 --no-cpu-baseline --distinct 512 > $OUT/${TAG}_bench_roundtrip48.json 2> $OUT/${TAG}_bench_roundtrip48.err; echo "rt exit $?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_roundtrip48.json"))
print("roundtrip48", round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms", d["roofline"].get("kernels_ms"))
PY
