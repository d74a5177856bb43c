#!/bin/bash
# validation of the self-resetting TNS list + the file16 workload
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/r2_v9_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2_v9_pytest.log
python bench.py > $OUT/r2_v9_bench_decode48.json 2> $OUT/r2_v9_bench_decode48.err; echo "bench exit $?"; cat $OUT/r2_v9_bench_decode48.json
python bench.py --workload file16 > $OUT/r2_v9_bench_file16.json 2> $OUT/r2_v9_bench_file16.err; echo "file16 exit $?"; cat $OUT/r2_v9_bench_file16.json; tail -5 $OUT/r2_v9_bench_file16.err
python bench.py --workload decode16 --streams 8192 --no-cpu-baseline > $OUT/r2_v9_bench_decode16_8k.json 2>/dev/null; cat $OUT/r2_v9_bench_decode16_8k.json
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?"
