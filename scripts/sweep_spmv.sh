#!/bin/bash
# scratch: blocked SpMV tunables (chunk non-zeros, ring depth, CTAs per SM)
for v in "2048 2 3" "1024 4 3" "1024 3 4" "1024 2 6" "1024 3 5" "512 4 6" "512 6 4"; do
  set -- $v
  export OB200_EXTRA_NVCC="-DOB200_BLK_CHUNK=$1 -DOB200_BLK_STAGES=$2 -DOB200_BLK_CTAS=$3"
  python -c "from oofem_b200 import build; build.build(force=True)" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "chunk=$1 stages=$2 ctas=$3: $(python scripts/time_spmv_modes.py 2>&1 | tail -1)"
done
