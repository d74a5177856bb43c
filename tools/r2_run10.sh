#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r2_i}
timeout 1200 python -m pytest tests/test_encoder_gpu.py tests/test_clip200_gpu.py tests/test_full_size_gpu.py -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${T}_pytest.log
python bench.py --workload roundtrip48 --steps 30 --no-cpu-baseline --distinct 256 > $OUT/${T}_bench_roundtrip48.json 2> $OUT/${T}_bench_roundtrip48.err
python - <<PY
import json
d = json.load(open("$OUT/${T}_bench_roundtrip48.json"))
print("roundtrip48 value", round(d["value"]/1e6,2), "ms/step", round(d["ms_per_step"],3), {k.split("::")[-1]: round(x,3) for k,x in d["roofline"]["kernels_ms"].items()})
PY
for o in 5 6 7; do
  LC3B_DQ_OCC=$o python bench.py --steps 200 --no-secondary --no-cpu-baseline > $OUT/${T}_bench_decode48_occ$o.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("$OUT/${T}_bench_decode48_occ$o.json"))
print("dequant occ=$o: decode48 value", round(d["value"]/1e6,2), "ms/step", round(d["ms_per_step"],4), {k.split("::")[-1]: round(x,4) for k,x in d["roofline"]["kernels_ms"].items()})
PY
done
