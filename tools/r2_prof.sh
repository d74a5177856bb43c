#!/bin/bash
# ncu --set full of the steady-state launches matching a kernel regex inside the default bench command; summaries only
# come back (the report stays on the box)
set -u
KRE=${1:-dequant}; TAG=${2:-r2_prof}; N=${3:-1}
OUT=gpurun_out; mkdir -p $OUT /tmp/ncu_reps
LC3B_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$KRE" -c $N -o /tmp/ncu_reps/$TAG \
    python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --no-secondary ${BENCH_ARGS:-} > $OUT/${TAG}.log 2>&1
python tools/ncu_summary.py /tmp/ncu_reps/$TAG.ncu-rep $OUT/${TAG}_ncu_summary.md $OUT/${TAG}_ncu_summary.json > /dev/null 2>&1 || echo "summary failed"
cat $OUT/${TAG}_ncu_summary.md
: > $OUT/${TAG}_source_hotspots.txt
for k in $(seq 0 $((N-1))); do python tools/ncu_lines.py /tmp/ncu_reps/$TAG.ncu-rep $k 40 >> $OUT/${TAG}_source_hotspots.txt 2>/dev/null; done
cat $OUT/${TAG}_source_hotspots.txt
for k in $(seq 0 $((N-1))); do python tools/ncu_sass.py /tmp/ncu_reps/$TAG.ncu-rep $k 25; done 2>&1 | tee $OUT/${TAG}_sass_hotspots.txt
