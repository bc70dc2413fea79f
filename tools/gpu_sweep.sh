#!/bin/bash
# usage: tools/gpu_sweep.sh <tag-for-logs>   (run on the GPU box through gpurun)
tag=$1
L=$PWD/2dto3dreconstruction_b200/lib
if [ -z "$NO_TESTS" ]; then timeout 400 python -m pytest tests/test_gpu_nn.py tests/test_gpu_icp.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -4; fi
SWEEP=${SWEEP_MAIN:-0:32:2.0,1:32:2.0} timeout 280 python tools/sweep_nn.py 8 30 1.0 > gpurun_out/sweep_$tag.log 2>&1
cat gpurun_out/sweep_$tag.log | tail -14
for v in $VARIANTS; do
  echo "== variant $v" | tee -a gpurun_out/sweep_$tag.log
  R3D_LIB_PATH=$L/libr3d_b200_$v.so SWEEP=${SWEEP_VAR:-1:32:2.0} timeout 200 python tools/sweep_nn.py 8 30 1.0 2>&1 | tail -6 | tee -a gpurun_out/sweep_$tag.log
done
echo "== stats" | tee -a gpurun_out/sweep_$tag.log
for it in ${STATS_ITERS:-1 30}; do echo iters $it; R3D_LIB_PATH=$L/libr3d_b200_stats.so SWEEP=${SWEEP_STATS:-1:32:2.0} timeout 100 python tools/sweep_nn.py 8 $it 1.0 2>&1 | tail -8; done 2>&1 | tee -a gpurun_out/sweep_$tag.log
