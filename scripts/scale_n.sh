#!/bin/bash
# usage: bash scripts/scale_n.sh "<N list>" -- strong and weak runs at several rank counts on one box
mkdir -p gpurun_out
for N in $1; do
  for mode in strong weak; do
    extra=""; [ $mode = strong ] && extra="--scaling strong --nx 256"
    tag=bench_n${N}_${mode}_r02
    if [ $N = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py"; fi
    timeout 600 $cmd --gpus $N --steps 10 --warmup 3 --no-parity --no-cpu-baseline $extra > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", "elements/s %.4g" % d["value"], "pcg it/s %.1f" % d["pcg_iters_per_s"], "ms_asm %.3f ms_pcg %.3f" % (d["ms_assembly"], d["ms_pcg"]), "e2e %.4g" % d["e2e"]["value"], d["clocks"]["reasons"])
except Exception as e:
    print("$tag parse failed", e); print(open("gpurun_out/$tag.err").read()[-800:])
PY
  done
done
