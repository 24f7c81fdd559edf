#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list, ncu full capture of the top kernels.
# Usage (under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; head -c 600 $OUT/bench_$TAG.json; echo; tail -3 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --cg-iters 20 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'spmv_block_kernel|lspace_cluster_kernel|cg_xr_p_kernel' -c 6 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 1 --cg-iters 2 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
echo "== BASELINE configs 3 (LTRSpace, one partition) and 4 (MisesMat NR) at full size"
timeout 400 python scripts/run_configs.py ltrspace mises > $OUT/configs_$TAG.jsonl 2> $OUT/configs_$TAG.err; echo "configs rc=$?"; cut -c1-400 $OUT/configs_$TAG.jsonl
if [ "$2" = "ref" ]; then
echo "== bench reference arm"
timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json | head -c 1500; echo
fi
ls -la $OUT | tail -12
