#!/bin/bash
# parity of the small-batch path + where its time goes (ncu launch lists are serialised, cold-cache timings: shares only)
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${1:-r2_c}
timeout 1500 python -m pytest tests -m gpu -q --durations=6 -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${T}_pytest.log
for w in decode16 mixed encode48; do
  python bench.py --workload $w --quick --steps 200 > $OUT/${T}_quick_$w.json 2> $OUT/${T}_quick_$w.err || tail -5 $OUT/${T}_quick_$w.err
  echo "$w: $(cat $OUT/${T}_quick_$w.json)"
done
LC3B_DEQUANT=thread python bench.py --workload decode16 --quick --steps 200 2>/dev/null | sed 's/^/decode16 thread-dequant: /'
LC3B_DEQUANT=thread python bench.py --workload mixed --quick --steps 200 2>/dev/null | sed 's/^/mixed thread-dequant: /'
for w in decode16 mixed; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${T}_launches_$w.csv \
      python bench.py --workload $w --quick --steps 3 --warmup 3 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/${T}_launches_$w.csv")) if len(r) > 5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
last = collections.OrderedDict()
for r in rows[-200:]:
    last.setdefault(r[4], []).append(float(r[-1].replace(",", "")))
print("$w kernels (last launches, us):", {k.split("(")[0][-40:]: round(sum(v) / len(v) / 1e3, 1) for k, v in last.items()})
PY
done
