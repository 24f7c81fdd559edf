#!/bin/bash
# scratch: rebuild the library with different gather tunables on the GPU box and time the assembly
for v in "24 16 3" "32 20 2" "32 16 2" "16 16 3" "24 12 3" "16 8 4"; do
  set -- $v
  export OB200_EXTRA_NVCC="-DOB200_GROUP_VISITS=$1 -DOB200_ROUND_ELEMS=$2 -DOB200_GATHER_CTAS=$3"
  python -c "from oofem_b200 import build; build.build(force=True)" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "variant visits=$1 round=$2 ctas=$3: $(python scripts/time_assembly.py 2>&1 | tail -1)"
done
