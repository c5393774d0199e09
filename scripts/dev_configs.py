"""Per-kernel times of BASELINE configs 3-5 (stand-in workloads, tests/workloads.py) on one GPU.
CFG=4|5|3  B=<hypotheses on this GPU>  ITERS=<n>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import scene_util as su, workloads as wl
from diffdope import _native as nat

cfg_id = int(os.environ.get("CFG", "5")); iters = int(os.environ.get("ITERS", "20"))
w = {4: wl.config4, 5: wl.config5, 3: wl.config3}[cfg_id]()
if cfg_id == 3:
    w = dict(w["objects"][3], P=w["P"], H=w["H"], W=w["W"], B=w["B"], losses=w["losses"], name=w["name"])
B = int(os.environ.get("B", str(w["B"])))
sc = nat.NativeScene(w["pos"], w["tri"], uv=w.get("uv"), tex=w.get("tex")) if w.get("tex") is not None else nat.NativeScene(w["pos"], w["tri"], vtx_color=w["vtx_color"])
sc.set_camera(w["P"], w["H"], w["W"])
out = sc.render(torch.from_numpy(w["q_gt"][None]).cuda(), torch.from_numpy(w["t_gt"][None]).cuda(), want=("rgb", "depth", "rast"))
cov = (out["rast"][0, ..., 3] > 0).float()
g = dict(rgb=out["rgb"][0].contiguous(), depth=(out["depth"][0] * cov).contiguous(), seg=cov.contiguous())
sc.set_target(g["rgb"], g["depth"], g["seg"])
L = w["losses"]
c = nat.make_loss_cfg(L.get("l1_rgb_with_mask", False), L.get("l1_depth_with_mask", False), L.get("l1_mask", False), L.get("weight_rgb", 1), L.get("weight_depth", 1),
                      L.get("weight_mask", 1), L.get("l1_edge", False) and not os.environ.get("NO_EDGE"), L.get("weight_edge", 1))
lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 2.0)).cuda()
sched = [20 * 0.1 ** (i / max(iters - 1, 1) + 1) for i in range(iters)]
def run():
    q = torch.from_numpy(np.tile(w["q0"], (B, 1))).cuda().contiguous(); t = torch.from_numpy(np.tile(w["t0"], (B, 1))).cuda().contiguous()
    return sc.optimize(q, t, lr, sched, c, keep_history=False)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
sc.profile_begin(); run(); k, n = sc.profile_end()
it = n["pixel_kernel"]
print(w["name"]); print("covered fraction %.3f, V=%d T=%d, %dx%d, B=%d, iters=%d" % (float(cov.mean()), sc.V, sc.T, w["W"], w["H"], B, iters))
print("per-iter us:", {a: round(1e3 * b / it, 1) for a, b in k.items()}, "| one call: %.3f ms/iter, %.0f hyp*iter/s" % (ms / iters, B * iters / (ms * 1e-3)))
