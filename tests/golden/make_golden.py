"""Generate tests/golden/example_q25.npz from the CPU oracle (oracle/refpath.py) on the reference's
example scene at quarter resolution. The reference itself cannot be run here (its hot path is
nvdiffrast, absent), so these are oracle outputs pinned as a regression fixture:
the CPU suite checks the oracle still reproduces them, the GPU suite checks the CUDA path against
them without needing the oracle's runtime.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import scene_util as su  # noqa: E402
from oracle import refpath  # noqa: E402

RESIZE, B, ITERS = 0.25, 2, 6
CFG = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)


def main():
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = {k: torch.from_numpy(v) for k, v in su.example_targets(RESIZE).items()}
    H, W = gt["rgb"].shape[:2]
    qs, ts = su.perturbed_poses(q, t, B)
    lr = su.lr_multipliers(B)
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    P = su.projection()
    logged, gq, gtr, r = refpath.forward_backward(mesh, P, qs, ts, gt, lr, CFG, H, W)
    rast = r["rast_out"].detach().numpy()
    ids = rast[..., 3].astype(np.int32)
    ys, xs = np.nonzero(ids.max(0) > 0)
    y0, y1, x0, x1 = ys.min() - 2, ys.max() + 3, xs.min() - 2, xs.max() + 3
    hyper = dict(nb_iterations=ITERS - 1, base_lr=20.0, lr_decay=0.1, learning_rate_base=1)
    # trajectory with small multipliers: the reference's L1 / sign-gradient SGD is chaotic for multipliers
    # >~ 1 (a 1e-7 change of the start pose moves the final pose by 0.1-0.8 mm on one and the same
    # implementation, scripts/dev_chaos.py), so a trajectory can only be pinned in the small-step regime
    opt_lr = np.array([0.1, 0.3], dtype=np.float32)
    opt = refpath.run_optimization(mesh, P, qs, ts, gt, opt_lr, CFG, hyper, H, W)
    np.savez_compressed(
        os.path.join(HERE, "example_q25.npz"),
        resize=RESIZE, H=H, W=W, quat=qs, trans=ts, lr=lr, bbox=np.array([y0, y1, x0, x1]),
        tri_id=ids,
        uvz=rast[:, y0:y1, x0:x1, :3],
        rgb=r["rgb"].detach().numpy()[:, y0:y1, x0:x1],
        depth=r["depth"].detach().numpy()[:, y0:y1, x0:x1],
        mask=r["mask"].detach().numpy()[:, y0:y1, x0:x1, 0],
        mtx=r["mtx"].detach().numpy(),
        loss=np.stack([logged["rgb"].numpy(), logged["depth"].numpy(), logged["mask_selection"].numpy()], 1),
        grad=np.concatenate([gq, gtr], 1),
        opt_lr=opt_lr, opt_poses=opt["poses"], opt_final=opt["final"],
        opt_losses=np.stack([opt["losses"]["rgb"], opt["losses"]["depth"], opt["losses"]["mask_selection"]], -1),
    )
    print("wrote example_q25.npz", H, W, "covered", int((ids > 0).sum()))


if __name__ == "__main__":
    main()
