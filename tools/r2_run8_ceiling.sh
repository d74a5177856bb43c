#!/bin/bash
# 8-GPU box: what it is, and what it can copy (no kernels)
set -u
OUT=gpurun_out; mkdir -p $OUT
( nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"; numactl -H 2>/dev/null; free -g | head -2; nvidia-smi topo -m; nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv ) > $OUT/r2_box8.txt 2>&1
python tools/host_copy_ceiling.py --reps 10 > $OUT/r2_ceiling8.log 2>&1; tail -40 $OUT/r2_ceiling8.log
cp $OUT/host_copy_ceiling.json $OUT/r2_host_copy_ceiling_8gpu.json
