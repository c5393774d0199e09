import sys, time, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/diff-dope_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
B=int(os.environ.get('B','64')); iters=int(os.environ.get('ITERS','50')); win=int(os.environ.get('WIN','640'))
arr=su.example_mesh_arrays(); q,t=su.example_pose(); P=su.projection_native()
gt=su.example_targets(1.0); H,W=gt['rgb'].shape[:2]
window=su.centred_window(gt['segmentation'],win,H,W); print('window',window)
sc=nat.NativeScene(arr['pos'],arr['tri'],arr['uv'],arr['tex']); sc.set_camera(P,H,W); sc.set_window(*window)
dev='cuda'
g={k:torch.from_numpy(v).to(dev) for k,v in gt.items()}
seg1=g['segmentation'][...,0].contiguous()
sc.set_target(g['rgb'],g['depth'],seg1)
lr=torch.from_numpy(su.lr_multipliers(B)).to(dev)
c=nat.make_loss_cfg(True,True,True,0.7,1.0,1.0)
sched=[20*0.1**(i/iters+1) for i in range(iters)]
def run():
    qd=torch.from_numpy(np.tile(q,(B,1))).to(dev).contiguous(); td=torch.from_numpy(np.tile(t,(B,1))).to(dev).contiguous()
    ph,lh=sc.optimize(qd,td,lr,sched,c)
    return qd,td,ph,lh
for _ in range(2): run()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); qd,td,ph,lh=run(); e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)
print('B',B,'iters',iters,'ms total',ms,'ms/iter',ms/iters,'h*it/s',B*iters/(ms*1e-3))
print('loss first',lh[0,0].cpu().numpy(),'last',lh[-1].mean(0).cpu().numpy())
print('pose0',ph[0,0].cpu().numpy(),'poseN',qd[0].cpu().numpy(),td[0].cpu().numpy())
