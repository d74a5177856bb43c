#!/bin/bash
# round 2, first GPU call: parity on the new 200-frame corpus, baseline numbers of the round-1 kernels on it, box facts
set -u
OUT=gpurun_out; mkdir -p $OUT
( nproc; lscpu | head -30; numactl -H 2>/dev/null; nvidia-smi topo -m; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv ) > $OUT/r2_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/r2_a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/r2_a_pytest.log
python bench.py > $OUT/r2_a_bench_decode48.json 2> $OUT/r2_a_bench_decode48.err; echo "bench rc=$?"; cat $OUT/r2_a_bench_decode48.json
python tools/host_copy_ceiling.py --reps 10 > $OUT/r2_a_ceiling1.log 2>&1; tail -30 $OUT/r2_a_ceiling1.log
