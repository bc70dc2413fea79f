#!/bin/bash
# usage: tools/gpu_ab.sh <tag> <variant> [<variant> ...]   (on the GPU box through gpurun)
# Per-launch ncu metrics (time, warp instructions, issue utilisation) of the 30 nn launches of one 8-pair chunk, for the product
# library ("main") and experimental twins (libr3d_b200_<variant>.so): what an A/B timing difference is made of.
tag=$1; shift
L=$PWD/2dto3dreconstruction_b200/lib
for v in "$@"; do
  lib=$L/libr3d_b200.so; [ "$v" != main ] && lib=$L/libr3d_b200_$v.so
  R3D_LIB_PATH=$lib WARM=0 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none --kernel-name-base demangled -k regex:nn_query --csv --log-file gpurun_out/ab_${tag}_$v.csv \
     python tools/profile_icp.py 8 ${PARTIAL:-1.0} 30 32 2.0 1 > gpurun_out/ab_${tag}_$v.log 2>&1
done
if [ -n "$FULL_IT" ]; then      # one --set full capture of ICP iteration $FULL_IT (1-based) with the product library
  WARM=0 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:nn_query -s $((FULL_IT-1)) -c 1 \
     -o gpurun_out/ab_${tag}_it$FULL_IT -f python tools/profile_icp.py 8 ${PARTIAL:-1.0} 30 32 2.0 1 > gpurun_out/ab_${tag}_full.log 2>&1
fi
