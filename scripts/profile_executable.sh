#!/bin/bash
# Where does the wall time of `oofem_cuda -f big.in` go?  Sampling profile (scripts/probes/sprof.c, LD_PRELOAD, SIGPROF at 200 Hz of
# process CPU time) of OOFEM's executable with the plugin on the generated input of scripts/e2e_executable.py.
# usage: scripts/profile_executable.sh [tag] [nx ny nz]
TAG=${1:-x}
NX=${2:-125}; NY=${3:-32}; NZ=${4:-32}
OUT=gpurun_out
mkdir -p $OUT /tmp/sprof_run
gcc -O2 -shared -fPIC -o /tmp/sprof_run/libsprof.so scripts/probes/sprof.c || exit 1
python - <<P
import sys
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
import e2e_executable as e
print(e.write_case('/tmp/sprof_run', ($NX, $NY, $NZ), '1e-3'))
P
cd /tmp/sprof_run
( time env OOFEM_B200_TIMING=1 SPROF_OUT=/tmp/sprof_run/sprof.out LD_PRELOAD=/tmp/sprof_run/libsprof.so timeout 300 $OLDPWD/plugin/_build/oofem_cuda -f cuda.in ) > $OLDPWD/$OUT/exe_prof_$TAG.log 2>&1
cd $OLDPWD
tail -12 $OUT/exe_prof_$TAG.log
python scripts/probes/sprof_sym.py $(ls -S /tmp/sprof_run/sprof.out.* | head -1) > $OUT/exe_profile_$TAG.txt 2>&1
head -150 $OUT/exe_profile_$TAG.txt
