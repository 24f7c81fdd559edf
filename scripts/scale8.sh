#!/bin/bash
# One 8-GPU box: weak scaling at 1M hex/GPU with the parity cases, BASELINE config 5 (4M hex/GPU = 32M hex), strong scaling.
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", "elements/s %.4g" % d["value"], "pcg it/s %.1f" % d["pcg_iters_per_s"], "ms_asm %.3f ms_pcg %.3f" % (d["ms_assembly"], d["ms_pcg"]), "e2e %.4g" % d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_sharers"), d["clocks"])
except Exception as e:
    print("$tag parse failed", e); print(open("gpurun_out/$tag.err").read()[-1500:])
PY
}
run bench_n8_weak_r02 --steps 10 --warmup 3
run bench_n8_config5_r02 --steps 5 --warmup 3 --nx 1000 --no-parity
run bench_n8_strong_r02 --steps 10 --warmup 3 --scaling strong --nx 256 --no-parity
