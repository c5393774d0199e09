"""GPU parity on BASELINE.json configs 3-5. The BOP assets are not in the tree: the workloads
are the labelled stand-ins of tests/workloads.py. Small sizes are compared with the oracle; full sizes are
checked through size-independent properties (shard equivalence, determinism, convergence to the pose the
target was rendered from)."""
import numpy as np
import pytest
import torch

import scene_util as su
import workloads as wl
from test_gpu_parity import _angle_deg, _cfg, _loss_table, _nat

pytestmark = pytest.mark.gpu


def _scene(n, w):
    if w.get("tex") is not None:
        sc = n.NativeScene(w["pos"], w["tri"], uv=w["uv"], tex=w["tex"])
    else:
        sc = n.NativeScene(w["pos"], w["tri"], vtx_color=w["vtx_color"])
    sc.set_camera(w["P"], w["H"], w["W"])
    return sc


def _oracle_mesh(w):
    from oracle import refpath

    return refpath.Mesh(w["pos"], w["tri"], w.get("uv"), w.get("tex"), w.get("vtx_color"))


def _oracle_targets(w, P, H, W):
    from oracle import refpath

    r = refpath.render(_oracle_mesh(w), P, torch.from_numpy(w["q_gt"][None]), torch.from_numpy(w["t_gt"][None]), H, W)
    return wl.targets_from_render(r["rgb"][0].numpy(), r["depth"][0].numpy(), r["rast_out"][0, ..., 3].numpy())


def _gpu_targets(sc, w):
    out = sc.render(torch.from_numpy(w["q_gt"][None]).cuda(), torch.from_numpy(w["t_gt"][None]).cuda(), want=("rgb", "depth", "rast"))
    cov = (out["rast"][0, ..., 3] > 0).float()
    return dict(rgb=out["rgb"][0].contiguous(), depth=(out["depth"][0] * cov).contiguous(), segmentation=cov.contiguous())


def _hyps(w, B, seed=3, rot=0.02, tr=0.01):
    rng = np.random.default_rng(seed)
    qs = np.tile(w["q0"], (B, 1)) + rng.normal(0, rot, (B, 4)).astype(np.float32)
    ts = np.tile(w["t0"], (B, 1)) + rng.normal(0, tr, (B, 3)).astype(np.float32)
    return qs.astype(np.float32), ts.astype(np.float32)


def _compare(sc, w, P, H, W, gt, qs, ts, lr, losses, n):
    from oracle import refpath

    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    loss, grad = sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(n, losses))
    logged, gq, gtr, r = refpath.forward_backward(_oracle_mesh(w), P, qs, ts, {k: torch.from_numpy(v) for k, v in gt.items()}, lr, losses, H, W)
    out = sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), want=("rast",))
    assert np.array_equal(r["rast_out"].detach().numpy()[..., 3], out["rast"].cpu().numpy()[..., 3]), "coverage / triangle ids bit-exact"
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, qs.shape[0]), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    su.record("compare_%dx%d_T%d_B%d" % (H, W, w["tri"].shape[0], qs.shape[0]), grad_rel_err=su.grad_rel_err(go, gg))
    assert np.abs(go).max() > 0 and np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
    return g


def test_config4_small_matches_oracle():
    n = _nat()
    w = wl.config4(0.25)
    sc = _scene(n, w)
    gt = _oracle_targets(w, w["P"], w["H"], w["W"])
    assert gt["segmentation"][..., 0].sum() > 400
    qs, ts = _hyps(w, 3)
    _compare(sc, w, w["P"], w["H"], w["W"], gt, qs, ts, su.lr_multipliers(3), w["losses"], n)


def test_config5_small_matches_oracle():
    n = _nat()
    w = wl.config5(0.125, tex_size=256)
    sc = _scene(n, w)
    gt = _oracle_targets(w, w["P"], w["H"], w["W"])
    assert 0.35 < gt["segmentation"][..., 0].mean() < 0.65, "the object fills about half of the window"
    qs, ts = _hyps(w, 2)
    _compare(sc, w, w["P"], w["H"], w["W"], gt, qs, ts, su.lr_multipliers(2), w["losses"], n)


def test_config3_small_matches_oracle():
    n = _nat()
    c = wl.config3(0.25)
    assert len(c["objects"]) == 8
    for k in (0, 6):  # the smallest-scale object and a cycled, shifted one
        w = dict(c["objects"][k], P=c["P"], H=c["H"], W=c["W"])
        sc = _scene(n, w)
        gt = _oracle_targets(w, c["P"], c["H"], c["W"])
        assert gt["segmentation"][..., 0].sum() > 50
        qs, ts = _hyps(w, 2, rot=0.01, tr=0.005)
        _compare(sc, w, c["P"], c["H"], c["W"], gt, qs, ts, su.lr_multipliers(2), c["losses"], n)


def _schedule(iters, base_lr=20.0, decay=0.1):
    nb = max(iters - 1, 1)
    return [base_lr * decay ** (it / nb + 1) for it in range(iters)]


def test_config4_full_size_properties():
    """256 hypotheses x 100 iterations at 720x540: two shards of 128 equal one batch of 256 bit for bit,
    the run is deterministic, and the best hypothesis ends closer to the rendered ground truth than it started."""
    n = _nat()
    w = wl.config4()
    sc = _scene(n, w)
    g = _gpu_targets(sc, w)
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    B, iters = w["B"], w["iters"]
    lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 3.0)).cuda()
    cfg = _cfg(n, w["losses"])
    sched = _schedule(iters)
    q0 = torch.from_numpy(np.tile(w["q0"], (B, 1))).cuda().contiguous()
    t0 = torch.from_numpy(np.tile(w["t0"], (B, 1))).cuda().contiguous()
    qa, ta = q0.clone(), t0.clone()
    pa, la = sc.optimize(qa, ta, lr, sched, cfg)
    qb, tb = q0.clone(), t0.clone()
    for lo in (0, 128):
        qs, ts = qb[lo:lo + 128].contiguous(), tb[lo:lo + 128].contiguous()
        ps, ls = sc.optimize(qs, ts, lr[lo:lo + 128].contiguous(), sched, cfg, b_global=B)
        assert torch.equal(ps, pa[:, lo:lo + 128]) and torch.equal(ls, la[:, lo:lo + 128])
        assert torch.equal(qs, qa[lo:lo + 128]) and torch.equal(ts, ta[lo:lo + 128])
    qc, tc = q0.clone(), t0.clone()
    pc, lc = sc.optimize(qc, tc, lr, sched, cfg)
    assert torch.equal(pc, pa) and torch.equal(lc, la)
    # convergence: argmin over hypotheses of the mean logged loss (DiffDope.get_argmin)
    best = int(la[-1, :, 1:3].mean(-1).argmin())
    fin_q, fin_t = qa[best].cpu().numpy(), ta[best].cpu().numpy()
    e0 = np.abs(w["t0"] - w["t_gt"]).max()
    e1 = np.abs(fin_t - w["t_gt"]).max()
    # the mask term has a floor (antialiased render vs binary segmentation on the silhouette ring); depth has none
    assert la[-1, best, 1] < 0.6 * la[0, best, 1], "depth loss of the selected hypothesis drops by 40 % or more"
    assert la[-1, best, 1:3].sum() < la[0, best, 1:3].sum()
    assert e1 < e0, "translation error shrinks (%.4f -> %.4f)" % (e0, e1)


def test_config5_full_size_properties():
    """1024 hypotheses at 1024^2, 50k triangles, full loss stack incl. Sobel edge: a few iterations, deterministic,
    finite, and the loss-ROI covers about half of the window (HBM-stress shape)."""
    n = _nat()
    w = wl.config5()
    sc = _scene(n, w)
    g = _gpu_targets(sc, w)
    assert 0.4 < float(g["segmentation"].mean()) < 0.6
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    B, iters = w["B"], 4
    lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 2.0)).cuda()
    cfg = _cfg(n, w["losses"])
    sched = _schedule(iters)
    res = []
    for _ in range(2):
        q = torch.from_numpy(np.tile(w["q0"], (B, 1))).cuda().contiguous()
        t = torch.from_numpy(np.tile(w["t0"], (B, 1))).cuda().contiguous()
        p, l = sc.optimize(q, t, lr, sched, cfg)
        res.append((q, t, p, l))
    assert all(torch.equal(a, b) for a, b in zip(res[0], res[1]))
    l = res[0][3]
    assert torch.isfinite(l).all() and (l[0] > 0).all()
    assert float(l[-1].sum(-1).min()) < float(l[0].sum(-1).min()), "some hypothesis improves within 4 iterations"
    # at the ground-truth pose every loss is (near) zero: the target is this renderer's own image
    q = torch.from_numpy(np.tile(w["q_gt"], (2, 1))).cuda().contiguous()
    t = torch.from_numpy(np.tile(w["t_gt"], (2, 1))).cuda().contiguous()
    loss, _ = sc.loss_grad(q, t, lr[:2].contiguous(), cfg)
    # (not exactly zero: the loss pass lets the compiler contract a*b+c, the image pass rounds every op separately)
    assert float(loss[:, :2].abs().max()) < 2e-6 and float(loss[:, 3].abs().max()) < 2e-6
    assert float(loss[:, 2].max()) < 5e-3  # antialiased mask vs binary segmentation: silhouette pixels only


def test_config1_window_single_hypothesis_matches_oracle():
    """BASELINE config 1: the example scene at full resolution, ONE hypothesis, 320x320 loss window centred
    on the segmentation, default losses (mask only) and the full stack: losses and gradient of that exact geometry against
    the oracle, then a short optimisation."""
    from oracle import refpath

    n = _nat()
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = su.example_targets(1.0)
    H, W = gt["rgb"].shape[:2]
    window = su.centred_window(gt["segmentation"], 320, H, W)
    assert abs(float(su.lr_multipliers(1)[0]) - 84.4437) < 1e-3  # the first uniform(0.01, 100) draw after random.seed(0)
    # that draw is meant for full-frame means: over a 320x320 window the mean (and the gradient) is 20x larger and the very
    # first step moves the object by ~11 units, out of the window, in the reference's own algebra. The step-by-step
    # comparison uses a multiplier from the non-expanding regime instead.
    lr = np.array([0.01], dtype=np.float32)
    sc = n.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(su.projection(), H, W)
    sc.set_window(*window)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    gt_t = {k: torch.from_numpy(v) for k, v in gt.items()}
    iters = 4
    sched = [refpath.lr_schedule(it, 49, 20.0, 0.1) for it in range(iters)]  # the first 4 of config 1's 50 iterations
    for losses in (dict(l1_mask=True, weight_mask=1.0), dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)):
        # one step: losses and gradient of this exact geometry (full resolution, window, one hypothesis)
        loss, grad = sc.loss_grad(torch.from_numpy(q[None]).cuda(), torch.from_numpy(t[None]).cuda(), torch.from_numpy(lr).cuda(), _cfg(n, losses))
        logged, gq, gtr, _ = refpath.forward_backward(mesh, su.projection(), q[None], t[None], gt_t, lr, losses, H, W, window=window)
        assert np.allclose(loss.cpu().numpy(), _loss_table(logged, 1), rtol=1e-4, atol=1e-10)
        go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
        su.record("config1_window_%d_losses" % len([k for k in losses if k.startswith("l1_")]), grad_rel_err=su.grad_rel_err(go, gg))
        assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
        # a few iterations run and move the pose (trajectories are compared with the oracle in test_gpu_parity.py)
        qd, td = torch.from_numpy(q[None]).cuda().contiguous(), torch.from_numpy(t[None]).cuda().contiguous()
        ph, lh = sc.optimize(qd, td, torch.from_numpy(lr).cuda(), sched, _cfg(n, losses))
        assert sc.last_launch_count() == 3 * iters + 1  # one hypothesis: no internal split
        assert np.allclose(lh[0].cpu().numpy(), _loss_table(logged, 1), rtol=1e-4, atol=1e-10)
        assert torch.isfinite(ph).all() and torch.isfinite(lh).all()
        assert not torch.equal(ph[0], ph[-1])
