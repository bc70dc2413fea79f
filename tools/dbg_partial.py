import sys, importlib, numpy as np, os
sys.path.insert(0,'/root/repo')
from oracle import ref_port as O
ops=importlib.import_module('2dto3dreconstruction_b200.ops'); C=importlib.import_module('2dto3dreconstruction_b200._capi')
a,b=O.synth_depth_pair(3); A0,B=O.depth_to_voxel_ld(a),O.depth_to_voxel_ld(b)
d=ops.pack_xyzw(C.to_device(B)); do=ops._offsets([len(B)])
np.set_printoptions(precision=5, suppress=True, linewidth=150)
for n in (4305, 20000, 70000):
    A=A0[:n]
    s=ops.pack_xyzw(C.to_device(A)); so=ops._offsets([len(A)])
    r2=ops.icp_batch(s,so,d,do,1,len(A),len(B),max_iterations=1,tolerance=0.0,partial_set_size=1.0, want_last=True)
    keep=r2['keep'].cpu().numpy().astype(bool); idx=r2['idx'].cpu().numpy()
    Tref=O.best_fit_transform(A[keep], B[idx[keep]])
    print(n, os.environ.get('R3D_FORCE_UNFUSED'), 'kept',keep.sum(), 'T err', np.abs(r2['T'][0].cpu().numpy()-Tref).max())
