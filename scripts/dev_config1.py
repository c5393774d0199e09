import os, sys
sys.path[:0] = ['/root/repo', '/root/repo/diff-dope_b200', '/root/repo/tests']
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
import bench
dev = torch.device('cuda', 0)
arr = su.example_mesh_arrays(); q, t = su.example_pose(); gt = su.example_targets(1.0); H, W = gt['rgb'].shape[:2]
sc = nat.NativeScene(arr['pos'], arr['tri'], arr['uv'], arr['tex']); sc.set_camera(su.projection_native(), H, W)
sc.set_window(*su.centred_window(gt['segmentation'], 320, H, W))
g = {k: torch.from_numpy(v).to(dev) for k, v in gt.items()}
sc.set_target(g['rgb'], g['depth'], g['segmentation'][..., 0].contiguous())
ALL = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
cfg = bench._loss_cfg(nat, ALL)
q0 = torch.from_numpy(np.tile(q, (1, 1))).to(dev).contiguous(); t0 = torch.from_numpy(np.tile(t, (1, 1))).to(dev).contiguous()
for name, lr, sched in (('bench lr 0.01 + bench schedule', torch.tensor([0.01], device=dev), bench.lr_schedule(50)),
                        ('multiplier of dev_time + bench schedule', torch.from_numpy(su.lr_multipliers(1)).to(dev), bench.lr_schedule(50)),
                        ('multiplier + dev_time schedule', torch.from_numpy(su.lr_multipliers(1)).to(dev), [20 * 0.1 ** (i / 50 + 1) for i in range(50)])):
    for hist in (False, True):
        def run():
            qq, tt = q0.clone(), t0.clone()
            sc.optimize(qq, tt, lr, sched, cfg, keep_history=hist)
        ms = bench.time_call(run, reps=5)
        sc.profile_begin(); run(); k, n = sc.profile_end()
        print(name, 'history', hist, ': %.1f us/iter (events around the call); kernels alone' % (1e3 * ms / 50), {a: round(1e3 * b / 50, 1) for a, b in k.items()})
