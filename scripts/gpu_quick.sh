#!/bin/bash
# quick GPU pass: parity tests + bench (no profiler)
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
