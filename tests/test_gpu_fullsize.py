"""GPU parity at the BASELINE.json configurations' REAL geometry, against the oracle:
config 2 = the bench workload itself (1920x1080 frame, 640x640 loss window, rgb + depth + mask, hypotheses of a
64-hypothesis job with the seeded multipliers), and one full-size hypothesis batch of each of configs 3, 4, 5
(stand-in workloads of tests/workloads.py). The oracle needs ~1 s per hypothesis-iteration at these sizes."""
import numpy as np
import pytest
import torch

import scene_util as su
import workloads as wl
from test_gpu_configs import _compare, _hyps, _oracle_targets, _scene
from test_gpu_parity import _angle_deg, _cfg, _loss_table, _nat

pytestmark = pytest.mark.gpu

ALL = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)


def test_config2_real_geometry_matches_oracle():
    """BASELINE configs[1] as bench.py runs it: full-resolution example scene, 640x640 window centred on the segmentation,
    rgb + depth + mask (0.7 / 1 / 1), four hypotheses of the 64-hypothesis job with their `random.seed(0)` multipliers
    (hypothesis 0: 84.4, and the three smallest draws 40: 0.124, 35: 1.41, 52: 8.05), B_global = 64 in the mean, distinct start
    poses. Triangle ids, barycentrics, z/w, rgb, depth bit-equal inside the window, losses and gradient 1e-4, then the first three
    SGD iterations of the 200-iteration schedule against `refpath.run_optimization`."""
    from oracle import refpath

    n = _nat()
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = su.example_targets(1.0)
    H, W = gt["rgb"].shape[:2]
    assert (H, W) == (1080, 1920)
    window = su.centred_window(gt["segmentation"], 640, H, W)
    y0, x0, wh, ww = window
    hyp, BG = [0, 40, 35, 52], 64
    B = len(hyp)
    lr = su.lr_multipliers(BG)[hyp].copy()
    assert abs(float(lr[0]) - 84.4437) < 1e-3 and abs(float(lr[1]) - 0.12427) < 1e-4 and abs(float(lr[3]) - 8.0538) < 1e-3
    qs, ts = su.perturbed_poses(q, t, B, seed=5, rot_deg=1.0, trans=0.01)
    P = su.projection()
    sc = n.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(P, H, W)
    sc.set_window(*window)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    gt_t = {k: torch.from_numpy(v) for k, v in gt.items()}
    qd, td, lrd = torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda()

    # forward: the window of the oracle's full-frame render
    out = sc.render(qd, td)
    logged, gq, gtr, r = refpath.forward_backward(mesh, P, qs, ts, gt_t, lr, ALL, H, W, window=window, b_global=BG)
    sl = (slice(None), slice(y0, y0 + wh), slice(x0, x0 + ww))
    rast_o = r["rast_out"].detach().numpy()[sl]
    rast_g = out["rast"].cpu().numpy()
    assert (rast_o[..., 3] > 0).sum() > B * 20000, "the object covers ~24k pixels per hypothesis at full resolution"
    assert np.array_equal(rast_o[..., 3], rast_g[..., 3]), "triangle ids / coverage bit-exact inside the window"
    assert np.array_equal(rast_o, rast_g), "barycentrics and z/w bit-exact"
    assert np.array_equal(r["rgb"].detach().numpy()[sl], out["rgb"].cpu().numpy())
    assert np.array_equal(r["depth"].detach().numpy()[sl], out["depth"].cpu().numpy())
    assert np.abs(r["mask"].detach().numpy()[sl][..., 0] - out["mask"].cpu().numpy()).max() <= 2e-7
    assert np.array_equal(r["mtx"].detach().numpy(), out["mtx"].cpu().numpy())

    # losses + gradient of the exact bench configuration
    loss, grad = sc.loss_grad(qd, td, lrd, _cfg(n, ALL), b_global=BG)
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    errs = [su.grad_rel_err(go[b], gg[b]) for b in range(B)]  # per hypothesis: the multipliers span three decades
    assert max(errs) <= 1e-4, "gradient rel err per hypothesis %r" % (errs,)

    # the first three iterations of the 200-iteration schedule
    iters = 3
    hyper = dict(nb_iterations=199, base_lr=20.0, lr_decay=0.1, learning_rate_base=1)
    sched = [refpath.lr_schedule(it, 199, 20.0, 0.1) for it in range(iters)]
    qo, to = qd.clone().contiguous(), td.clone().contiguous()
    ph, lh = sc.optimize(qo, to, lrd, sched, _cfg(n, ALL), b_global=BG)
    ref = refpath.run_optimization(mesh, P, qs, ts, gt_t, lr, ALL, hyper, H, W, window=window, b_global=BG, stop_after=iters)
    ph, lh = ph.cpu().numpy(), lh.cpu().numpy()
    assert np.array_equal(ph[0], ref["poses"][0])
    fin = np.concatenate([qo.cpu().numpy(), to.cpu().numpy()], 1)
    lo = np.stack([np.stack([ref["losses"][k][it] for k in ("rgb", "depth", "mask_selection")], 1) for it in range(iters)])
    meas = dict(grad_rel_err=errs, pose_diff=[[float(np.abs(ph[it, b] - ref["poses"][it, b]).max()) for b in range(B)] for it in range(iters)],
                loss_rel=[[float(np.abs(lh[it, b, :3] / lo[it, b] - 1).max()) for b in range(B)] for it in range(iters)],
                final_trans_diff=[float(np.abs(fin[b, 4:] - ref["final"][b, 4:]).max()) for b in range(B)],
                final_angle_deg=[float(_angle_deg(fin[b, :4], ref["final"][b, :4])) for b in range(B)],
                step0_move=[float(np.abs(ref["poses"][1, b] - ref["poses"][0, b]).max()) for b in range(B)])
    # how far the ORACLE ITSELF drifts from itself when hypothesis 0's start quaternion moves by one float32 ulp: the yardstick
    # for what any other implementation of the same algebra can be expected to reproduce in this regime
    q1 = qs[:1].copy()
    q1[0, 0] = np.nextafter(q1[0, 0], np.float32(2.0))
    ref1 = refpath.run_optimization(mesh, P, q1, ts[:1], gt_t, lr[:1], ALL, hyper, H, W, window=window, b_global=BG, stop_after=iters)
    meas["oracle_self_drift_1ulp_hyp0"] = [float(np.abs(ref1["poses"][it, 0] - ref["poses"][it, 0]).max()) for it in range(iters)] + [float(np.abs(ref1["final"][0] - ref["final"][0]).max())]
    su.record("config2_real_geometry", **{k: np.asarray(v) for k, v in meas.items()})
    # Iteration 0 is the single-step comparison and holds for every hypothesis. Afterwards: multiplier 84.4 with B_global = 64 is the
    # expanding regime -- its first step moves the object by 13 mm and 4 degrees, one silhouette sample
    # flipping is 5e-4 of the mask loss, and a pair with a short edge can carry 1 % of the gradient -- so hypothesis 0 is held to
    # "same pose entering iteration 1" only; the three hypotheses with multipliers <= 8 must track the oracle through all three steps.
    for b in range(B):
        assert np.allclose(lh[0, b, :3], lo[0, b], rtol=1e-4, atol=1e-9), meas
        assert meas["pose_diff"][1][b] <= 1e-6, meas
    for b in (1, 2, 3):
        for it in (1, 2):
            assert meas["pose_diff"][it][b] <= 1e-5 and meas["loss_rel"][it][b] <= 5e-4, (b, it, meas)
        # north_star bar on the pose after the three steps: 0.1 mm (1 unit = 100 mm at scale 0.01) and 0.1 degree
        assert meas["final_trans_diff"][b] < 1e-3 and meas["final_angle_deg"][b] < 0.1, meas


@pytest.mark.parametrize("which", ["config3", "config4", "config5"])
def test_full_size_hypotheses_match_oracle(which):
    """One full-size batch of each stand-in workload against the oracle: ids bit-exact, losses and gradient 1e-4."""
    n = _nat()
    if which == "config3":
        c = wl.config3()
        w = dict(c["objects"][3], P=c["P"], H=c["H"], W=c["W"], losses=c["losses"])
        B = 2
    elif which == "config4":
        w, B = wl.config4(), 3
    else:
        w, B = wl.config5(), 1
    assert (w["H"], w["W"]) == {"config3": (480, 640), "config4": (540, 720), "config5": (1024, 1024)}[which]
    sc = _scene(n, w)
    gt = _oracle_targets(w, w["P"], w["H"], w["W"])
    qs, ts = _hyps(w, B, rot=0.01, tr=0.005)
    _compare(sc, w, w["P"], w["H"], w["W"], gt, qs, ts, su.lr_multipliers(B), w["losses"], n)
