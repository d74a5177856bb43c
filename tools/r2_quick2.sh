#!/bin/bash
# decoder parity tests (incl. full size) + A/B of LC3B_SPLIT on the decode48 line + secondary block
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_decoder_gpu.py tests/test_clip200_gpu.py tests/test_multi_frame_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > $OUT/r2_q9_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r2_q9_pytest.log
bash tools/r2_ab.sh r2_ab_split LC3B_SPLIT=1 LC3B_SPLIT=2 LC3B_SPLIT=4
for s in 65536 131072; do BENCH_ARGS="--streams $s" bash tools/r2_ab.sh r2_ab_split_$s LC3B_SPLIT=1 LC3B_SPLIT=2 LC3B_SPLIT=4; done
