#!/bin/bash
# final bench lines of the round (default command with its secondary block, round trip, both file workloads, mixed) + launch list
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=r2_final
python bench.py > $OUT/${TAG}_bench_decode48.json 2> $OUT/${TAG}_bench_decode48.err || echo "bench decode48 failed"
for w in roundtrip48 file48 file16 mixed decode16; do
  python bench.py --workload $w --distinct 512 > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err || echo "bench $w failed"
done
python - <<PY
import json
for w in ("decode48", "roundtrip48", "file48", "file16", "mixed", "decode16"):
    d = json.load(open("$OUT/${TAG}_bench_%s.json" % w))
    print(w, round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]/1e6, 2), "frac", round(d["roofline"]["frac"], 4), "launches", d.get("gpu_launches"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
d = json.load(open("$OUT/${TAG}_bench_decode48.json"))
for k, v in d.get("secondary", {}).items(): print(" ", k, round(v["ms_per_step"], 4), "ms", round(v["value"]/1e6, 2), "M")
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_decode48.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/${TAG}_launches_decode48.log 2>&1
