#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r2_d}
timeout 1500 python -m pytest tests -m gpu -q --durations=6 -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${T}_pytest.log
for w in decode16 mixed; do
  python bench.py --workload $w --quick --steps 200 2>/dev/null | sed "s/^/$w: /"
  LC3B_DEQUANT=thread python bench.py --workload $w --quick --steps 200 2>/dev/null | sed "s/^/$w thread-dequant: /"
done
python bench.py --steps 100 --no-secondary > $OUT/${T}_bench_decode48.json 2> $OUT/${T}_bench_decode48.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/${T}_bench_decode48.json"))
print("decode48 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["kernels_ms"])
PY
