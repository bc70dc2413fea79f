import sys, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo')
from oracle import ref_port as O
pipeline=importlib.import_module('2dto3dreconstruction_b200.pipeline'); icp=importlib.import_module('2dto3dreconstruction_b200.utils.icp')
fa,fb=[],[]
for k in range(5):
    a,b=O.synth_depth_pair(20+k,h=120,w=160); fa.append(a); fb.append(b)
big_a=np.zeros((5,480,640),np.uint16); big_b=np.zeros((5,480,640),np.uint16)
big_a[:,180:300,240:400]=np.stack(fa); big_b[:,180:300,240:400]=np.stack(fb)
big_a[3]=0
kw=dict(max_iterations=8,tolerance=0.0,partial_set_size=1.0,pairs_per_chunk=2,n_streams=2)
mode=sys.argv[1]
if mode=='dev_first':
    dev=pipeline.register_depth_pairs(torch.from_numpy(big_a).cuda(),torch.from_numpy(big_b).cuda(),**kw)
    Td=dev['T'].cpu().numpy()
host=pipeline.register_depth_pairs(big_a,big_b,**kw)
if mode!='dev_first': Td=host['T']
print('dev==host',np.array_equal(Td,host['T']))
for p in (0,1,2,4):
    A,B=O.depth_to_voxel_ld(big_a[p]),O.depth_to_voxel_ld(big_b[p])
    T,_,it=icp.icp(A,B,max_iterations=8,tolerance=0.0)
    print(p,'diff',np.abs(T-Td[p]).max(), 'iters', it, int(host['iters'][p]))
