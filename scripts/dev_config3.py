"""BASELINE config 3 (stand-in, tests/workloads.py): 8 objects x 128 hypotheses x 100 iterations, full loss stack incl.
the Sobel-edge extension, one GPU. Objects refined one after the other (the reference's BOP loop) vs concurrently
(one stream per object, what diffdope.run_optimization_batched does)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import scene_util as su, workloads as wl
from diffdope import _native as nat

c3 = wl.config3()
B, iters = int(os.environ.get("B", c3["B"])), int(os.environ.get("ITERS", c3["iters"]))
L = c3["losses"]
cfg = nat.make_loss_cfg(True, True, True, L["weight_rgb"], L["weight_depth"], L["weight_mask"], not os.environ.get("NO_EDGE"), L["weight_edge"])
objs = []
for o in c3["objects"]:
    sc = nat.NativeScene(o["pos"], o["tri"], uv=o["uv"], tex=o["tex"])
    sc.set_camera(c3["P"], c3["H"], c3["W"])
    out = sc.render(torch.from_numpy(o["q_gt"][None]).cuda(), torch.from_numpy(o["t_gt"][None]).cuda(), want=("rgb", "depth", "rast"))
    cov = (out["rast"][0, ..., 3] > 0).float()
    g = (out["rgb"][0].contiguous(), (out["depth"][0] * cov).contiguous(), cov.contiguous())
    sc.set_target(*g)
    objs.append((sc, o, g))
lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 2.0)).cuda()
sched = [20 * 0.1 ** (i / max(iters - 1, 1) + 1) for i in range(iters)]
streams = [torch.cuda.Stream() for _ in objs]

def start(o):
    return (torch.from_numpy(np.tile(o["q0"], (B, 1))).cuda().contiguous(), torch.from_numpy(np.tile(o["t0"], (B, 1))).cuda().contiguous())

def sequential():
    res = []
    for sc, o, _ in objs:
        q, t = start(o)
        sc.optimize(q, t, lr, sched, cfg, keep_history=False)
        res.append((q, t))
    return res

def batched():
    cur = torch.cuda.current_stream()
    res = []
    for (sc, o, _), st in zip(objs, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            q, t = start(o)
            sc.optimize(q, t, lr, sched, cfg, keep_history=False)
            res.append((q, t))
    for st in streams:
        cur.wait_stream(st)
    return res

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r

ms_s, rs = timed(sequential)
ms_b, rb = timed(batched)
same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(rs, rb))
n = len(objs) * B * iters
print(c3["name"])
print("%d objects x %d hypotheses x %d iterations, %dx%d, edge loss %s" % (len(objs), B, iters, c3["W"], c3["H"], "off" if os.environ.get("NO_EDGE") else "on"))
print("sequential: %.1f ms (%.0f hyp*iter/s)   concurrent streams: %.1f ms (%.0f hyp*iter/s)   identical results: %s" % (ms_s, n / ms_s * 1e3, ms_b, n / ms_b * 1e3, same))
err = [float(np.abs(r[1][0].cpu().numpy() - o["t_gt"]).max()) for r, (_, o, _) in zip(rb, objs)]
print("translation error of hypothesis 0 per object after refinement (units):", ["%.4f" % e for e in err], " start:", ["%.4f" % float(np.abs(o["t0"] - o["t_gt"]).max()) for _, o, _ in objs])
