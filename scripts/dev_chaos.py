"""How chaotic is the reference's SGD? Perturb the start pose by 1e-7 and watch the GPU trajectories separate."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/diff-dope_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
arr=su.example_mesh_arrays(); q,t=su.example_pose(); P=su.projection_native()
gt=su.example_targets(0.5); H,W=gt['rgb'].shape[:2]
sc=nat.NativeScene(arr['pos'],arr['tri'],arr['uv'],arr['tex']); sc.set_camera(P,H,W)
g={k:torch.from_numpy(v).cuda() for k,v in gt.items()}
sc.set_target(g['rgb'],g['depth'],g['segmentation'])
lrs=[0.1,0.3,1,3,10,30,84]
B=len(lrs)
for name,c in (('mask',nat.make_loss_cfg(False,False,True)),('all',nat.make_loss_cfg(True,True,True,0.7,1,1))):
  for iters in (12,61):
    sched=[20*0.1**(i/(iters-1)+1) for i in range(iters)]
    res=[]
    for eps in (0.0,1e-7):
        qd=torch.from_numpy(np.tile(q,(B,1))).cuda().contiguous(); td=torch.from_numpy(np.tile(t,(B,1))).cuda().contiguous()
        td[:,0]+=eps
        ph,lh=sc.optimize(qd,td,torch.tensor(lrs).float().cuda(),sched,c)
        res.append((torch.cat([qd,td],1).cpu().numpy(), lh.cpu().numpy()))
    d=np.abs(res[0][0]-res[1][0]).max(1)
    moved=np.abs(res[0][0]-np.concatenate([q,t])[None]).max(1)
    print(name,'iters',iters,'final-pose sep per lr',dict(zip(lrs,np.round(d,6))),'moved',np.round(moved,4),'loss0->N',res[0][1][0,:,:].sum(-1)[0],np.round(res[0][1][-1].sum(-1),5))
