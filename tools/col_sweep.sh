#!/bin/bash
# column-sweep shape / stage / prefetch experiments at 100^3: tools/col_sweep.sh out.txt
out=$1; shift
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python tools/col_check.py time ${DIMS:-100 100 100} 2>&1 | grep -E "apply|column sweeps|rror" >> $out; }
for shape in 8x4/1x3 8x4/3x1 8x4/1x2 8x4/2x1 8x4/2x2 4x8/3x1 4x8/1x3 8x4/1x4 8x4/2x3 8x4/1x1; do
  run OPMGPU_DEBUG=1 OPMGPU_COL_SHAPE=$shape
done
run OPMGPU_COL_SHAPE=8x4/1x3 OPMGPU_COL_PF=0
run OPMGPU_COL_SHAPE=8x4/1x3 OPMGPU_COL_PF=24
run OPMGPU_COL_SHAPE=8x4/1x3 OPMGPU_COL_STAGES=3
run OPMGPU_COL_SHAPE=8x4/1x3 OPMGPU_COL_STAGES=4
cat $out
