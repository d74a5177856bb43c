#!/bin/bash
# quick check of an encoder kernel change: encoder parity tests + the roundtrip48 line
set -u
TAG=${1:-r2_qe}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_encoder_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
python bench.py --workload roundtrip48 --no-cpu-baseline --distinct 512 > $OUT/${TAG}_bench_roundtrip48.json 2> $OUT/${TAG}_bench_roundtrip48.err; echo "rt exit $?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_roundtrip48.json"))
print("roundtrip48", round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms", {k.split("::")[-1]: round(v, 3) for k, v in d["roofline"].get("kernels_ms", {}).items()})
PY
python bench.py --workload encode48 --no-cpu-baseline --distinct 512 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('encode48', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],4))"
