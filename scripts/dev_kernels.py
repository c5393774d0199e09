"""Per-kernel ms/iteration of the bench workload (library event hook)."""
import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/diff-dope_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
B=int(os.environ.get('B','64')); iters=int(os.environ.get('ITERS','50')); win=int(os.environ.get('WIN','640')); res=float(os.environ.get('RES','1.0'))
arr=su.example_mesh_arrays(); q,t=su.example_pose(); P=su.projection_native()
gt=su.example_targets(res); H,W=gt['rgb'].shape[:2]
window=su.centred_window(gt['segmentation'],win,H,W)
sc=nat.NativeScene(arr['pos'],arr['tri'],arr['uv'],arr['tex']); sc.set_camera(P,H,W); sc.set_window(*window)
g={k:torch.from_numpy(v).cuda() for k,v in gt.items()}
sc.set_target(g['rgb'],g['depth'],g['segmentation'][...,0].contiguous())
lr=torch.from_numpy(su.lr_multipliers(B)).cuda()
c=nat.make_loss_cfg(True,True,True,0.7,1.0,1.0)
sched=[20*0.1**(i/iters+1) for i in range(iters)]
def run(hist=False):
    qd=torch.from_numpy(np.tile(q,(B,1))).cuda().contiguous(); td=torch.from_numpy(np.tile(t,(B,1))).cuda().contiguous()
    return sc.optimize(qd,td,lr,sched,c,keep_history=hist)
run(); run(); torch.cuda.synchronize()
sc.profile_begin(); run(); k,n=sc.profile_end()
it=n['pixel_kernel']
print(os.environ.get('TAG',''), 'per-iter us:', {a:round(1e3*b/it,1) for a,b in k.items()}, 'total', round(1e3*sum(k.values())/it,1))
