#!/bin/bash
# quick check of a decoder kernel change: decoder parity tests + the decode48 line without the secondary block
set -u
TAG=${1:-r2_q}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_decoder_gpu.py tests/test_clip200_gpu.py tests/test_multi_frame_gpu.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
python bench.py --no-cpu-baseline --no-secondary ${BENCH_ARGS:-} > $OUT/${TAG}_bench_decode48.json 2> $OUT/${TAG}_bench_decode48.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_decode48.json"))
print("decode48", round(d["value"]/1e6, 2), "M", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"]/1e6, 2), d["roofline"]["kernels_ms"])
PY
