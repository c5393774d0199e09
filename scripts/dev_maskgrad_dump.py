"""Parity investigation: per-tile gradient partials of the kernel vs the oracle's per-pair terms binned by tile
(config-1 geometry, mask loss only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
from oracle import refpath, nvdr

arr = su.example_mesh_arrays(); q, t = su.example_pose(); gt = su.example_targets(1.0)
H, W = gt["rgb"].shape[:2]
window = su.centred_window(gt["segmentation"], 320, H, W)
lr = np.array([0.01], dtype=np.float32)
sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
sc.set_camera(su.projection(), H, W); sc.set_window(*window)
g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
sc.set_target(g["rgb"], g["depth"], g["segmentation"])
ALL = os.environ.get("LOSSES", "mask") == "all"
cfg = nat.make_loss_cfg(ALL, ALL, True, 0.7, 1.0, 1.0)
loss, grad = sc.loss_grad(torch.from_numpy(q[None]).cuda(), torch.from_numpy(t[None]).cuda(), torch.from_numpy(lr).cuda(), cfg)
hyp = sc.debug_read(1, 208).view(np.int32)
rx0, ry0, rx1, ry1, tiles_x, tiles_y, tile_base = [int(v) for v in hyp[40:47]]
print("roi", rx0, ry0, rx1, ry1, "tiles", tiles_x, tiles_y, tile_base)
part = sc.debug_read(0, tiles_x * tiles_y * 80).view(np.float32).reshape(-1, 20)
out = sc.render(torch.from_numpy(q[None]).cuda(), torch.from_numpy(t[None]).cuda(), want=("mask", "rast"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "maskgrad_dump_%s.npz" % ("all" if ALL else "mask")), part=part, hyp=hyp, grad=grad.cpu().numpy(), loss=loss.cpu().numpy(),
         mask=out["mask"].cpu().numpy(), ids=out["rast"].cpu().numpy()[..., 3], window=np.array(window))
print("kernel grad", grad.cpu().numpy())
