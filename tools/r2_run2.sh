#!/bin/bash
# round 2, second GPU call: parity of the plan/graph executors, the mixed-rate and sharded handles; small-batch numbers
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/r2_b_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 $OUT/r2_b_pytest.log
for g in 0 1; do
  for w in decode16 mixed encode48; do
    LC3B_GRAPH=$g python bench.py --workload $w --quick --steps 200 > $OUT/r2_b_quick_${w}_g$g.json 2> $OUT/r2_b_quick_${w}_g$g.err || tail -5 $OUT/r2_b_quick_${w}_g$g.err
    echo "graph=$g $w: $(cat $OUT/r2_b_quick_${w}_g$g.json)"
  done
done
python bench.py --steps 100 > $OUT/r2_b_bench_decode48.json 2> $OUT/r2_b_bench_decode48.err; echo "bench rc=$?"; tail -c 1500 $OUT/r2_b_bench_decode48.json
