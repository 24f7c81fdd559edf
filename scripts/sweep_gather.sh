#!/bin/bash
# scratch: rebuild the library with different gather tunables on the GPU box and time the assembly
for v in 2 8 6; do
  export OB200_EXTRA_NVCC="-DOB200_EL_PAD=$v"
  python -c "from oofem_b200 import build; build.build(force=True)" > /dev/null 2>&1 || { echo "build failed $v"; continue; }
  echo "el_pad=$v: $(python scripts/time_assembly.py 2>&1 | tail -1)"
done
