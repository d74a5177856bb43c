#!/bin/bash
# 8-GPU box: the sharded front end end to end (one process), and strong scaling of BASELINE config 5 under torchrun
set -u
OUT=gpurun_out; mkdir -p $OUT; T=r2_8gpu
python tools/bench_sharded.py --steps 30 > $OUT/${T}_sharded.log 2>&1; tail -5 $OUT/${T}_sharded.log | cut -c1-400
cp $OUT/bench_sharded.json $OUT/${T}_bench_sharded.json
P=29500
for n in 2 4 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 100 --warmup 5 --total-streams 262144 --no-secondary \
      > $OUT/${T}_strong${n}_decode48.json 2> $OUT/${T}_strong${n}_decode48.err
  python - <<PY
import json
d = json.load(open("$OUT/${T}_strong${n}_decode48.json"))
print("strong N=$n decode48: value", round(d["value"]/1e6,1), "M/s  ms/step", round(d["ms_per_step"],4), " e2e", round(d["e2e"]["value"]/1e6,1), "scaling", d["scaling"], d["config"].get("total_streams"))
PY
done
for n in 1 8; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 50 --warmup 5 --workload roundtrip48 --total-streams 262144 --distinct 256 --no-secondary \
      > $OUT/${T}_strong${n}_roundtrip48.json 2> $OUT/${T}_strong${n}_roundtrip48.err
  python - <<PY
import json
d = json.load(open("$OUT/${T}_strong${n}_roundtrip48.json"))
print("strong N=$n roundtrip48: value", round(d["value"]/1e6,2), "M/s  ms/step", round(d["ms_per_step"],4), " e2e", round(d["e2e"]["value"]/1e6,2))
PY
done
