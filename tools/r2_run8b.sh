#!/bin/bash
# 8-GPU box, final kernels: strong scaling of BASELINE config 5 under torchrun (decode48 at 1/2/4/8, round trip at 8)
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r2_8gpu_final
P=29600
for n in 1 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 100 --warmup 5 --total-streams 262144 --no-secondary --no-cpu-baseline \
      > $OUT/${T}_strong${n}_decode48.json 2> $OUT/${T}_strong${n}_decode48.err
  python - <<PY
import json
d = json.load(open("$OUT/${T}_strong${n}_decode48.json"))
print("strong N=$n decode48: value", round(d["value"]/1e6,1), "M/s  ms/step", round(d["ms_per_step"],4), " e2e", round(d["e2e"]["value"]/1e6,1))
PY
done
P=$((P+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 50 --warmup 5 --workload roundtrip48 --total-streams 262144 --distinct 256 --no-secondary --no-cpu-baseline \
    > $OUT/${T}_strong8_roundtrip48.json 2> $OUT/${T}_strong8_roundtrip48.err
python - <<PY
import json
d = json.load(open("$OUT/${T}_strong8_roundtrip48.json"))
print("strong N=8 roundtrip48: value", round(d["value"]/1e6,2), "M/s  ms/step", round(d["ms_per_step"],4), " e2e", round(d["e2e"]["value"]/1e6,2))
PY
