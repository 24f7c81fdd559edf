#!/bin/bash
# BASELINE configs[2]: one ~4M-element LTRSpace mesh (112x77x77 cells x 6 tetrahedra) element-partitioned over N GPUs of one box.
# usage: bash scripts/config3.sh N [extra bench flags]
N=$1; shift
mkdir -p gpurun_out
tag=bench_tet_n${N}_r02
if [ $N = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py"; fi
timeout 600 $cmd --gpus $N --etype ltrspace --scaling strong --nx 112 --ny 77 --nz 77 --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
echo "$tag rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", "elements/s %.4g" % d["value"], "pcg it/s %.1f" % d["pcg_iters_per_s"], "ms_asm %.3f ms_pcg %.3f" % (d["ms_assembly"], d["ms_pcg"]), "e2e %.4g" % d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_sharers"), d["roofline_assembly"]["kernel"], d["clocks"]["reasons"])
except Exception as e:
    print("$tag parse failed", e); print(open("gpurun_out/$tag.err").read()[-1500:])
PY
