for v in "" _t256b6 _t256b7 _t128b6 _t64b8 _t512b3; do
  R3D_LIB_PATH=$PWD/2dto3dreconstruction_b200/lib/libr3d_b200$v.so python tools/bench_kernels.py 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin)['K4B_ransac3d_sweep'];print('$v', round(d['ms'],3), round(d['frac_of_fp64_nominal'],3))"
done
