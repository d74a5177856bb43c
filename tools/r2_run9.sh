#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r2_h}
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_clip200_gpu.py tests/test_multi_frame_gpu.py tests/test_full_size_gpu.py tests/test_loss_harness.py -m gpu -q -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${T}_pytest.log
python bench.py --steps 200 --no-secondary --no-cpu-baseline > $OUT/${T}_bench_decode48.json 2> $OUT/${T}_bench_decode48.err
python - <<PY
import json
d = json.load(open("$OUT/${T}_bench_decode48.json"))
print("decode48 value", round(d["value"]/1e6,2), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e6,2), {k.split("::")[-1]: round(x,4) for k,x in d["roofline"]["kernels_ms"].items()})
PY
LC3B_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -o $OUT/${T}_dec48 \
    python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline > $OUT/${T}_dec48.log 2>&1
python tools/ncu_summary.py $OUT/${T}_dec48.ncu-rep $OUT/${T}_dec48_ncu_summary.md $OUT/${T}_dec48_ncu_summary.json | cut -c1-250
: > $OUT/${T}_dec48_source_hotspots.txt
for k in 0 1 2; do python tools/ncu_lines.py $OUT/${T}_dec48.ncu-rep $k 14 >> $OUT/${T}_dec48_source_hotspots.txt 2>/dev/null; echo >> $OUT/${T}_dec48_source_hotspots.txt; done
rm -f $OUT/${T}_dec48.ncu-rep
