#!/bin/bash
# Scratch: rebuild assemble_cluster.cu with different warp / ring configurations on the GPU box and time the kernel.
# usage: bash scripts/sweep_cluster.sh "C G R H [extra nvcc flags]" ...
cd $(dirname $0)/..
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for cfg in "$@"; do
  set -- $cfg
  nvcc -c ${SRC:-oofem_b200/csrc/assemble_cluster.cu} -x cu -I oofem_b200/csrc -o oofem_b200/csrc/assemble_cluster.o $FL -DOB200_CL_CWARPS=$1 -DOB200_CL_GWARPS=$2 -DOB200_CL_RECSLOTS=$3 -DOB200_CL_HSLOTS=$4 $5 $6 $7 -Xptxas -v 2>&1 | grep -A2 "lspace_cluster_kernelILb0" | grep -E "spill|Used" ; test -f oofem_b200/csrc/assemble_cluster.o || continue
  nvcc -shared -o oofem_b200/liboofem_b200.so oofem_b200/csrc/*.o -lcudart -ldl
  echo "== C=$1 G=$2 R=$3 H=$4 $5 $6 $7: $(timeout 120 python scripts/time_assembly2.py ${NXYZ:-250 64 64} 2>&1 | grep -E "^cluster|rror|relerr|CLPROF" | head -5)"
done
