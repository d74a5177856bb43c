#!/bin/bash
# Run on the GPU box (gpurun): every bench workload, the ncu launch list of the default bench command, and one
# `ncu --set full` capture of all kernels.  Outputs land in gpurun_out/ and are summarised into profiles/ afterwards
# with tools/ncu_summary.py and tools/ncu_lines.py.
set -u
TAG=${1:-r1_final}
OUT=gpurun_out
mkdir -p $OUT
for w in decode48 encode48 decode16 mixed roundtrip48 file48; do
  python bench.py --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err || echo "bench $w failed"
  tail -c 300 $OUT/${TAG}_bench_$w.json | head -c 300; echo
done
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_decode48.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_decode48.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches_decode48.log 2>&1
# one full step of the round trip = 8 encoder + 4 decoder kernels; 3 warm-up steps are skipped
ncu --set full --clock-control none --import-source on -k regex:"enc_|entropy|dequant|synth|ltpf_kernel" -s 36 -c 12 -o $OUT/${TAG}_all \
    python bench.py --workload roundtrip48 --steps 2 --warmup 3 --quick --no-cpu-baseline > $OUT/${TAG}_all.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"multi|plc_scan" -s 12 -c 4 -o $OUT/${TAG}_multi \
    python bench.py --workload file48 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_multi.log 2>&1
ls -la $OUT | tail -20
