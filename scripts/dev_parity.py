import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/diff-dope_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import scene_util as su
from oracle import nvdr, refpath
from diffdope import _native as nat
arr=su.example_mesh_arrays(); q,t=su.example_pose(); P=su.projection()
resize=0.5
gt=su.example_targets(resize); H,W=gt['rgb'].shape[:2]
B=3
qs,ts=su.perturbed_poses(q,t,B)
lr=su.lr_multipliers(B)
mesh=refpath.Mesh(arr['pos'],arr['tri'],arr['uv'],arr['tex'])
gt_t={k:torch.from_numpy(v) for k,v in gt.items()}
cfg=dict(l1_rgb_with_mask=True,weight_rgb=0.7,l1_depth_with_mask=True,weight_depth=1.0,l1_mask=True,weight_mask=1.0)
t0=time.time(); logged,gq,gtr,r=refpath.forward_backward(mesh,P,qs,ts,gt_t,lr,cfg,H,W); print('oracle s',time.time()-t0)
sc=nat.NativeScene(arr['pos'],arr['tri'],arr['uv'],arr['tex'])
sc.set_camera(P,H,W)
dev='cuda'
g={k:torch.from_numpy(v).to(dev) for k,v in gt.items()}
sc.set_target(g['rgb'],g['depth'],g['segmentation'])
qd=torch.from_numpy(qs).to(dev); td=torch.from_numpy(ts).to(dev)
out=sc.render(qd,td); torch.cuda.synchronize()
rast_o=r['rast_out'].detach().numpy(); rast_g=out['rast'].cpu().numpy()
print('tri id mismatches', (rast_o[...,3]!=rast_g[...,3]).sum(), 'covered', (rast_o[...,3]>0).sum())
print('coverage mismatches', ((rast_o[...,3]>0)!=(rast_g[...,3]>0)).sum())
print('rast uvz maxdiff', np.abs(rast_o[...,:3]-rast_g[...,:3]).max())
for k in ('rgb','depth'):
    a=r[k].detach().numpy(); b=out[k].cpu().numpy(); print(k,'maxabs',np.abs(a-b).max(),'maxrel',(np.abs(a-b)/(np.abs(a)+1e-6)).max())
a=r['mask'].detach().numpy()[...,0]; b=out['mask'].cpu().numpy(); print('mask maxabs',np.abs(a-b).max(), 'n diff', (a!=b).sum(), 'aa px', ((a>0)&(a<1)).sum())
print('mtx diff', np.abs(r['mtx'].detach().numpy()-out['mtx'].cpu().numpy()).max())
c=nat.make_loss_cfg(True,True,True,0.7,1.0,1.0)
loss,grad=sc.loss_grad(qd,td,torch.from_numpy(lr).to(dev),c); torch.cuda.synchronize()
print('loss gpu',loss.cpu().numpy()); print('loss ora',np.stack([logged['rgb'].numpy(),logged['depth'].numpy(),logged['mask_selection'].numpy()],1))
print('grad gpu',grad.cpu().numpy()); print('grad ora',np.concatenate([gq,gtr],1))
go=np.concatenate([gq,gtr],1); gg=grad.cpu().numpy()
print('grad rel err', np.abs(go-gg).max()/np.abs(go).max())
# --- autograd path: render_texture_batch + torch losses vs oracle gradient
import diffdope as dd
qt=torch.from_numpy(qs).to(dev).requires_grad_(True); tt=torch.from_numpy(ts).to(dev).requires_grad_(True)
qn=qt/torch.norm(qt,dim=1).reshape(-1,1)
mtx=dd.matrix_batch_44_from_position_quat(qn,tt)
proj=torch.from_numpy(P.astype(np.float32)).to(dev)
rr=dd.render_texture_batch(None,proj,mtx,torch.from_numpy(arr['pos']).to(dev),torch.from_numpy(arr['tri']).to(dev),[H,W],uv=torch.from_numpy(arr['uv']).to(dev),tex=torch.from_numpy(arr['tex']).to(dev))
lrt=torch.from_numpy(lr).to(dev)
seg=g['segmentation'][None]
l=(torch.mean(torch.abs((rr['rgb']-g['rgb'][None])*seg),(1,2,3))*lrt).mean()*0.7
l=l+(torch.mean(torch.abs((rr['depth']-g['depth'][None])*seg[...,0]),(1,2))*lrt).mean()
l=l+(torch.mean(torch.abs(rr['mask']-seg),(1,2,3))*lrt).mean()
l.backward()
ga=torch.cat([qt.grad,tt.grad],1).cpu().numpy()
print('autograd-path grad rel err', np.abs(go-ga).max()/np.abs(go).max())
