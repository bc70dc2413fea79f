L=$PWD/2dto3dreconstruction_b200/lib
for v in "" _st128 "" _st128; do R3D_LIB_PATH=$L/libr3d_b200$v.so python tools/bench_k1.py 2>&1 | tail -1; done
timeout 300 python -m pytest tests/test_gpu_backproject.py tests/test_gpu_nn.py -x -q 2>&1 | tail -2
