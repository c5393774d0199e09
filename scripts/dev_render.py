"""Image output (ddope_render) of the bench workload: 64 hypotheses, 640x640 window, rgb + depth + mask; CUDA-event time per call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import scene_util as su
from diffdope import _native as nat
B = int(os.environ.get("B", 64)); reps = int(os.environ.get("REPS", 10))
arr = su.example_mesh_arrays(); q, t = su.example_pose(); gt = su.example_targets(1.0)
H, W = gt["rgb"].shape[:2]
sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"]); sc.set_camera(su.projection_native(), H, W)
sc.set_window(*su.centred_window(gt["segmentation"], 640, H, W))
qd = torch.from_numpy(np.tile(q, (B, 1))).cuda().contiguous(); td = torch.from_numpy(np.tile(t, (B, 1))).cuda().contiguous()
for _ in range(3): sc.render(qd, td, want=("rgb", "depth", "mask"))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): sc.render(qd, td, want=("rgb", "depth", "mask"))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("render B=%d: %.4f ms per call, %.2f TB/s on 20 B/px" % (B, ms, B * 640 * 640 * 20 / ms / 1e9))
