"""Does splitting the hypotheses over concurrent streams overlap the (issue-bound) raster kernel of one part with the
(latency-bound) pixel kernel of another? NPARTS scenes of B/NPARTS hypotheses each, one stream per scene."""
import sys, os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
B = int(os.environ.get("B", "64")); iters = int(os.environ.get("ITERS", "100")); nparts = int(os.environ.get("NPARTS", "2"))
arr = su.example_mesh_arrays(); q, t = su.example_pose(); P = su.projection_native()
gt = su.example_targets(1.0); H, W = gt["rgb"].shape[:2]
window = su.centred_window(gt["segmentation"], 640, H, W)
g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
seg1 = g["segmentation"][..., 0].contiguous()
scs = []
for _ in range(nparts):
    sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"]); sc.set_camera(P, H, W); sc.set_window(*window)
    sc.set_target(g["rgb"], g["depth"], seg1); scs.append(sc)
lr = torch.from_numpy(su.lr_multipliers(B)).cuda()
c = nat.make_loss_cfg(True, True, True, 0.7, 1.0, 1.0)
sched = [20 * 0.1 ** (i / iters + 1) for i in range(iters)]
streams = [torch.cuda.Stream() for _ in range(nparts)]
per = B // nparts
def run():
    cur = torch.cuda.current_stream()
    outs = []
    for k, (sc, st) in enumerate(zip(scs, streams)):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            qd = torch.from_numpy(np.tile(q, (per, 1))).cuda().contiguous(); td = torch.from_numpy(np.tile(t, (per, 1))).cuda().contiguous()
            sc.optimize(qd, td, lr[k * per:(k + 1) * per].contiguous(), sched, c, b_global=B, keep_history=False)
            outs.append((qd, td))
    for st in streams: cur.wait_stream(st)
    return outs
run(); run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); o = run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("parts %d x %d hyp, pixel CTAs/SM %s: %.1f us/iter, %.0f hyp*iter/s  (q0 %s)" % (nparts, per, os.environ.get("DDOPE_PIXEL_CTAS_PER_SM", "4"), 1e3 * ms / iters, B * iters / (ms * 1e-3), o[0][0][0].cpu().numpy()))
