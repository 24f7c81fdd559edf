#!/bin/bash
# Scratch: A/B of the distributed CG tail (cooperative one-launch tail vs the six-launch sequence) under torchrun.
# usage: bash scripts/dist_ab.sh N "<bench args>"
N=$1; shift
for coop in 1 0; do
  OB200_CG_COOP=$coop timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+coop)) bench.py --gpus $N --steps 5 --warmup 3 --no-parity $@ 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('coop=$coop', '$@', 'pcg it/s', round(d['pcg_iters_per_s'],1), 'ms_pcg', round(d['ms_pcg'],3), 'elements/s', round(d['value']/1e6,1), 'M', 'launches', d['gpu_launches'])"
done
