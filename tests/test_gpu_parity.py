"""GPU (-m gpu): the CUDA hot path, called through the C ABI (ctypes -> libddope_b200.so), against
the CPU oracle on the same seeded inputs, against the committed golden fixture, and -- at the
benchmark's full size -- through size-independent properties.

Bars (BASELINE.json north_star): coverage / triangle ids bit-exact; rgb, depth, mask, loss values
within 1e-4 relative (they are in fact bit-equal for rgb/depth/barycentrics); gradients within
1e-4 of the largest component; final pose within 0.1 deg / 0.1 mm (0.001 scene units)."""
import os

import numpy as np
import pytest
import torch

import scene_util as su

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "example_q25.npz")
ALL = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)


def _nat():
    from diffdope import _native

    return _native


def _cfg(n, d):
    return n.make_loss_cfg(d.get("l1_rgb_with_mask", False), d.get("l1_depth_with_mask", False), d.get("l1_mask", False),
                           d.get("weight_rgb", 1.0), d.get("weight_depth", 1.0), d.get("weight_mask", 1.0),
                           d.get("l1_edge", False), d.get("weight_edge", 1.0))


def _loss_table(logged, B):
    z = np.zeros(B, np.float32)
    return np.stack([logged["rgb"].numpy() if "rgb" in logged else z, logged["depth"].numpy() if "depth" in logged else z,
                     logged["mask_selection"].numpy() if "mask_selection" in logged else z,
                     logged["edge"].numpy() if "edge" in logged else z], 1)


def _angle_deg(qa, qb):
    qa = qa / np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=-1, keepdims=True)
    d = np.clip(np.abs((qa * qb).sum(-1)), 0, 1)
    return np.degrees(2 * np.arccos(d))


class Example:
    def __init__(self, resize, window=None, seg1=False):
        n = _nat()
        self.n = n
        self.arr = su.example_mesh_arrays()
        self.q, self.t = su.example_pose()
        self.gt = su.example_targets(resize)
        self.H, self.W = self.gt["rgb"].shape[:2]
        self.P = su.projection()
        self.sc = n.NativeScene(self.arr["pos"], self.arr["tri"], self.arr["uv"], self.arr["tex"])
        self.sc.set_camera(self.P, self.H, self.W)
        self.g = {k: torch.from_numpy(v).cuda() for k, v in self.gt.items()}
        seg = self.g["segmentation"][..., 0].contiguous() if seg1 else self.g["segmentation"]
        self.sc.set_target(self.g["rgb"], self.g["depth"], seg)
        self.window = window
        if window is not None:
            self.sc.set_window(*window)

    def oracle_mesh(self):
        from oracle import refpath

        return refpath.Mesh(self.arr["pos"], self.arr["tri"], self.arr["uv"], self.arr["tex"])

    def gt_t(self):
        return {k: torch.from_numpy(v) for k, v in self.gt.items()}


@pytest.fixture(scope="module")
def ex_half():
    return Example(0.5)


def test_render_matches_oracle_bit_exact(ex_half):
    from oracle import refpath

    ex = ex_half
    B = 3
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    out = ex.sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
    r = refpath.render(ex.oracle_mesh(), ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W)
    ro, rg = r["rast_out"].numpy(), out["rast"].cpu().numpy()
    assert (ro[..., 3] > 0).sum() > 15000
    assert np.array_equal(ro[..., 3], rg[..., 3]), "triangle ids / coverage must be bit-exact"
    assert np.array_equal(ro[..., :3], rg[..., :3]), "barycentrics and z/w are computed with the same rounded ops"
    assert np.array_equal(r["mtx"].numpy(), out["mtx"].cpu().numpy())
    assert np.array_equal(r["rgb"].numpy(), out["rgb"].cpu().numpy())
    assert np.array_equal(r["depth"].numpy(), out["depth"].cpu().numpy())
    mo, mg = r["mask"].numpy()[..., 0], out["mask"].cpu().numpy()
    assert ((mo > 0) & (mo < 1)).sum() > 300, "antialiased silhouette pixels present"
    # the oracle interpolates the constant 1 (1 - u - v + u + v, off by <= 1 ulp); the kernel uses exactly 1
    assert np.abs(mo - mg).max() <= 1.2e-7


@pytest.mark.parametrize("losses", [ALL, dict(l1_mask=True, weight_mask=1.0), dict(l1_depth_with_mask=True, weight_depth=1.0),
                                    dict(l1_rgb_with_mask=True, weight_rgb=0.7)])
def test_loss_and_gradient_match_oracle(ex_half, losses):
    from oracle import refpath

    ex = ex_half
    B = 3
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = su.lr_multipliers(B)
    loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, losses))
    logged, gq, gtr, _ = refpath.forward_backward(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, losses, ex.H, ex.W)
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()


def test_window_and_shard_divisor(ex_half):
    """Loss window = slice of the full-frame problem; B_global keeps the reference's 1/B scale."""
    from oracle import refpath

    ex = ex_half
    win = su.centred_window(ex.gt["segmentation"], 128, ex.H, ex.W)
    win = (win[0] + 9, win[1] - 17, 100, 90)  # cut through the object so window edges matter
    ex.sc.set_window(*win)
    try:
        B = 2
        qs, ts = su.perturbed_poses(ex.q, ex.t, B)
        lr = su.lr_multipliers(B)
        loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, ALL), b_global=8)
        logged, gq, gtr, _ = refpath.forward_backward(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, ALL, ex.H, ex.W, window=win, b_global=8)
        assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
        go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
        assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
        out = ex.sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
        assert tuple(out["rgb"].shape) == (B, win[2], win[3], 3)
        r = refpath.render(ex.oracle_mesh(), ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W)
        y0, x0, h, w = win
        assert np.array_equal(r["rast_out"].numpy()[:, y0:y0 + h, x0:x0 + w, 3], out["rast"].cpu().numpy()[..., 3])
        assert np.abs(r["mask"].numpy()[:, y0:y0 + h, x0:x0 + w, 0] - out["mask"].cpu().numpy()).max() <= 1.2e-7
    finally:
        ex.sc.set_window(0, 0, ex.H, ex.W)


@pytest.mark.parametrize("hw,edge", [((7, 45), False), ((16, 32), False), ((17, 33), False), ((49, 95), False), ((33, 64), True), ((31, 97), True)])
def test_windows_that_do_not_fit_the_tile_grid(ex_half, hw, edge):
    """The pixel pass works on 32x16 tiles (32x32 with the edge loss), 4 pixel rows per thread. Windows smaller than a tile, one row /
    column larger than a multiple of it, and cut through the object at odd offsets: losses and gradients against the full-frame
    oracle, image output bit-equal to the oracle's slice (ragged last tile row / column, tiles with a single live warp)."""
    from oracle import refpath

    ex = ex_half
    c = su.centred_window(ex.gt["segmentation"], 128, ex.H, ex.W)  # 128 x 128 around the object's centre
    win = (c[0] + 64 - hw[0] // 2 + 5, c[1] + 64 - hw[1] // 2 - 7, hw[0], hw[1])
    losses = dict(ALL, l1_edge=True, weight_edge=0.5) if edge else ALL
    ex.sc.set_window(*win)
    try:
        B = 2
        qs, ts = su.perturbed_poses(ex.q, ex.t, B)
        lr = su.lr_multipliers(B)
        loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, losses))
        logged, gq, gtr, _ = refpath.forward_backward(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, losses, ex.H, ex.W, window=win)
        # atol: a window inside the object has a mask loss of exactly 0 here and of ~4e-9 in the oracle, which interpolates the constant 1
        # as u + v + (1 - u - v) = 1 +- 1 ulp (deliberate deviations); every other entry is O(0.1)
        assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-7)
        go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
        assert np.abs(go).max() > 0 and np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
        out = ex.sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
        r = refpath.render(ex.oracle_mesh(), ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W)
        y0, x0, h, w = win
        assert np.array_equal(r["rast_out"].numpy()[:, y0:y0 + h, x0:x0 + w, 3], out["rast"].cpu().numpy()[..., 3])
        assert np.array_equal(r["rgb"].numpy()[:, y0:y0 + h, x0:x0 + w], out["rgb"].cpu().numpy())
        assert np.array_equal(r["depth"].numpy()[:, y0:y0 + h, x0:x0 + w], out["depth"].cpu().numpy())
        assert np.abs(r["mask"].numpy()[:, y0:y0 + h, x0:x0 + w, 0] - out["mask"].cpu().numpy()).max() <= 1.2e-7
    finally:
        ex.sc.set_window(0, 0, ex.H, ex.W)


def test_golden_fixture():
    g = np.load(GOLDEN)
    ex = Example(float(g["resize"]))
    qd, td = torch.from_numpy(g["quat"]).cuda(), torch.from_numpy(g["trans"]).cuda()
    out = ex.sc.render(qd, td)
    assert np.array_equal(out["rast"].cpu().numpy()[..., 3].astype(np.int32), g["tri_id"])
    y0, y1, x0, x1 = g["bbox"]
    assert np.array_equal(out["rast"].cpu().numpy()[:, y0:y1, x0:x1, :3], g["uvz"])
    assert np.array_equal(out["rgb"].cpu().numpy()[:, y0:y1, x0:x1], g["rgb"])
    assert np.array_equal(out["depth"].cpu().numpy()[:, y0:y1, x0:x1], g["depth"])
    assert np.abs(out["mask"].cpu().numpy()[:, y0:y1, x0:x1] - g["mask"]).max() <= 1.2e-7
    loss, grad = ex.sc.loss_grad(qd, td, torch.from_numpy(g["lr"]).cuda(), _cfg(ex.n, ALL))
    assert np.allclose(loss.cpu().numpy()[:, :3], g["loss"], rtol=1e-4) and not loss[:, 3].any()
    assert np.abs(grad.cpu().numpy() - g["grad"]).max() <= 1e-4 * np.abs(g["grad"]).max()
    # 6 SGD iterations against the oracle's trajectory
    n = g["opt_poses"].shape[0]
    sched = [20.0 * 0.1 ** (it / (n - 1) + 1) for it in range(n)]
    ph, lh = ex.sc.optimize(qd.clone(), td.clone(), torch.from_numpy(g["opt_lr"]).cuda(), sched, _cfg(ex.n, ALL))
    assert np.allclose(lh.cpu().numpy()[..., :3], g["opt_losses"], rtol=5e-4, atol=1e-9)
    assert np.abs(ph.cpu().numpy() - g["opt_poses"]).max() < 1e-4


def test_optimisation_trajectory_final_pose(ex_half):
    """Reference loop (diffdope.py:1634-1714) for 12 iterations, default config losses (mask only)
    and the full stack: final pose within 0.1 deg / 0.1 mm of the oracle's."""
    from oracle import refpath

    ex = ex_half
    B, iters = 2, 12
    qs, ts = np.tile(ex.q, (B, 1)), np.tile(ex.t, (B, 1))
    # small multipliers from the reference's [0.01, 100] range. The L1 / sign-gradient SGD is chaotic for
    # multipliers >~ 1: a 1e-7 change of the start pose moves the final pose by 0.1-0.8 mm on one and the
    # same implementation (scripts/dev_chaos.py; the reference is not run-to-run reproducible there
    # either: unordered float atomics). The 0.1 deg / 0.1 mm bar is meaningful only
    # where the iteration is not expanding.
    lr = np.array([0.1, 0.3], dtype=np.float32)
    for losses in (dict(l1_mask=True, weight_mask=1.0), ALL):
        hyper = dict(nb_iterations=iters - 1, base_lr=20.0, lr_decay=0.1, learning_rate_base=1)
        o = refpath.run_optimization(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, losses, hyper, ex.H, ex.W)
        sched = [refpath.lr_schedule(it, iters - 1, 20.0, 0.1) for it in range(iters)]
        qd, td = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
        ph, lh = ex.sc.optimize(qd, td, torch.from_numpy(lr).cuda(), sched, _cfg(ex.n, losses))
        fin = np.concatenate([qd.cpu().numpy(), td.cpu().numpy()], 1)
        moved = np.abs(o["final"] - o["poses"][0]).max()
        assert moved > 2e-4, "the optimisation must actually move the pose"
        assert _angle_deg(fin[:, :4], o["final"][:, :4]).max() < 0.1
        assert np.abs(fin[:, 4:] - o["final"][:, 4:]).max() < 1e-3  # 0.1 mm = 0.001 units (scale 0.01 of mm)
        key = {"mask_selection": 2, "rgb": 0, "depth": 1}
        # logged losses: tight while no coverage decision has flipped (first iterations); afterwards a 1-ulp pose
        # difference may flip a pixel centre in or out of a triangle, which moves a mean-over-pixels loss by up to
        # 1/(H*W) per flipped pixel -- allow four such flips over the trajectory
        flip = 4.0 / (ex.H * ex.W)
        for k, v in o["losses"].items():
            ours = lh.cpu().numpy()[:, :, key[k]]
            assert np.allclose(ours[:3], v[:3], rtol=5e-4, atol=1e-9)
            assert np.allclose(ours, v, rtol=5e-4, atol=flip)


def _cube_scene(H=96, W=128, big=False):
    n = _nat()
    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32) * (0.9 if big else 0.5)
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    col = np.random.default_rng(3).random((8, 3)).astype(np.float32)
    from oracle import refpath

    P = refpath.projection_matrix(120.0, 120.0, W / 2 - 3.3, H / 2 + 1.7, W, H)
    sc = n.NativeScene(v, f, vtx_color=col)
    sc.set_camera(P, H, W)
    mesh = refpath.Mesh(v, f, vtx_color=col)
    rng = np.random.default_rng(1)
    gt = dict(rgb=rng.random((H, W, 3)).astype(np.float32), depth=(3 + rng.random((H, W))).astype(np.float32),
              segmentation=np.repeat((rng.random((H, W, 1)) > 0.4).astype(np.float32), 3, axis=2))
    return n, sc, mesh, P, gt, H, W


@pytest.mark.parametrize("big", [False, True])
def test_untextured_large_triangles(big):
    """Vertex-colour mesh (reference branch diffdope.py:229-231) with triangles far larger than the
    small-triangle fast path (64 px), partly outside the frame when `big`."""
    from oracle import refpath

    n, sc, mesh, P, gt, H, W = _cube_scene(big=big)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    qs = np.array([[0.3, 0.2, 0.1, 0.9], [0.0, 0.7, 0.1, 0.6]], dtype=np.float32)
    ts = np.array([[0.1, -0.05, -4.0], [-0.3, 0.2, -2.2 if big else -3.0]], dtype=np.float32)
    lr = np.array([1.0, 3.0], dtype=np.float32)
    out = sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
    r = refpath.render(mesh, P, torch.from_numpy(qs), torch.from_numpy(ts), H, W)
    assert np.array_equal(r["rast_out"].numpy(), out["rast"].cpu().numpy())
    assert np.array_equal(r["rgb"].numpy(), out["rgb"].cpu().numpy())
    assert np.abs(r["mask"].numpy()[..., 0] - out["mask"].cpu().numpy()).max() <= 1.2e-7
    loss, grad = sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(n, ALL))
    logged, gq, gtr, _ = refpath.forward_backward(mesh, P, qs, ts, {k: torch.from_numpy(v) for k, v in gt.items()}, lr, ALL, H, W)
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, 2), rtol=1e-4)
    go = np.concatenate([gq, gtr], 1)
    assert np.abs(go - grad.cpu().numpy()).max() <= 1e-4 * np.abs(go).max()


def test_triangles_crossing_the_camera_plane_match_oracle():
    """Near-plane handling: triangles with vertices behind the camera plane (w <= 0) are rasterised in homogeneous coordinates
    instead of being dropped (GL / nvdiffrast clip them). A ground plane passing under the camera, a slanted
    triangle with one vertex behind, one entirely behind, and a cube the camera sits inside of: ids / barycentrics / z/w bit-equal
    to the oracle, losses and gradients 1e-4, in both rasterisation modes."""
    from oracle import refpath

    n = _nat()
    v = np.array([[-2.0, -0.3, 1.0], [2.0, -0.3, 1.0], [2.0, -0.3, -6.0], [-2.0, -0.3, -6.0],
                  [0.3, 0.5, -2.0], [0.9, 0.1, -1.5], [0.5, 0.9, 0.7],
                  [-1.0, 0.2, 0.5], [-0.5, 0.3, 2.0], [-0.8, 0.9, 1.0]], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.int32)
    col = np.random.default_rng(5).random((10, 3)).astype(np.float32)
    H, W = 96, 128
    P = refpath.projection_matrix(90.0, 95.0, W / 2 - 3.3, H / 2 + 1.7, W, H)
    rng = np.random.default_rng(1)
    gt = dict(rgb=rng.random((H, W, 3)).astype(np.float32), depth=(1 + 3 * rng.random((H, W))).astype(np.float32),
              segmentation=np.repeat((rng.random((H, W, 1)) > 0.4).astype(np.float32), 3, axis=2))
    qs = np.array([[0.0, 0.0, 0.0, 1.0], [0.05, -0.08, 0.03, 0.99]], dtype=np.float32)
    ts = np.array([[0.02, -0.03, 0.0], [-0.05, 0.04, 0.1]], dtype=np.float32)
    lr = np.array([1.0, 2.0], dtype=np.float32)
    cube_v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32) * 0.9
    cube_f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    cube_col = np.random.default_rng(3).random((8, 3)).astype(np.float32)
    cube_q = np.array([[0.3, 0.2, 0.1, 0.9], [0.0, 0.7, 0.1, 0.6]], dtype=np.float32)
    cube_t = np.array([[0.2, 0.1, -0.6], [-0.1, 0.2, 0.3]], dtype=np.float32)  # the camera is inside the cube: every face crosses the camera plane or is behind it
    for vv, ff, cc, q, t in ((v, f, col, qs, ts), (cube_v, cube_f, cube_col, cube_q, cube_t)):
        sc = n.NativeScene(vv, ff, vtx_color=cc)
        sc.set_camera(P, H, W)
        g = {k: torch.from_numpy(x).cuda() for k, x in gt.items()}
        sc.set_target(g["rgb"], g["depth"], g["segmentation"])
        mesh = refpath.Mesh(vv, ff, vtx_color=cc)
        r = refpath.render(mesh, P, torch.from_numpy(q), torch.from_numpy(t), H, W)
        out = sc.render(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda())
        ro = r["rast_out"].numpy()
        assert (ro[..., 3] > 0).sum() > 2000
        assert np.array_equal(ro, out["rast"].cpu().numpy())
        assert np.array_equal(r["rgb"].numpy(), out["rgb"].cpu().numpy()) and np.array_equal(r["depth"].numpy(), out["depth"].cpu().numpy())
        assert np.abs(r["mask"].numpy()[..., 0] - out["mask"].cpu().numpy()).max() <= 2e-7
        logged, gq, gtr, _ = refpath.forward_backward(mesh, P, q, t, {k: torch.from_numpy(x) for k, x in gt.items()}, lr, ALL, H, W)
        go = np.concatenate([gq, gtr], 1)
        res = []
        for mode in ("zbuffer", "binned"):
            sc.set_raster_mode(mode)
            loss, grad = sc.loss_grad(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(lr).cuda(), _cfg(n, ALL))
            assert np.allclose(loss.cpu().numpy(), _loss_table(logged, 2), rtol=1e-4)
            assert np.abs(go - grad.cpu().numpy()).max() <= 1e-4 * np.abs(go).max()
            res.append((loss, grad))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def test_edge_cases_empty_coverage_and_errors():
    n, sc, mesh, P, gt, H, W = _cube_scene()
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    # object behind the camera / far outside the frame: nothing covered, finite losses, zero pose-rotation gradient
    qs = np.array([[0, 0, 0, 1.0], [0, 0, 0, 1.0]], dtype=np.float32)
    ts = np.array([[0, 0, 5.0], [40.0, 0, -4.0]], dtype=np.float32)
    out = sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
    assert out["rast"].abs().max().item() == 0 and out["mask"].max().item() == 0
    assert torch.allclose(out["depth"][0], torch.full_like(out["depth"][0], -5.0))  # background depth = -t_z
    loss, grad = sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.ones(2).cuda(), _cfg(n, ALL))
    assert torch.isfinite(loss).all() and torch.isfinite(grad).all()
    assert grad[:, :4].abs().max().item() == 0
    assert grad[0, 6].item() != 0  # background depth still pulls on t_z inside the segmentation
    with pytest.raises(RuntimeError):
        sc.set_window(0, 0, H + 1, W)
    with pytest.raises(RuntimeError):
        n.NativeScene(np.zeros((3, 3), np.float32), np.array([[0, 1, 5]], np.int32), vtx_color=np.zeros((3, 3), np.float32))
    sc2 = n.NativeScene(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.int32), vtx_color=np.zeros((3, 3), np.float32))
    with pytest.raises(RuntimeError):
        sc2.render(torch.zeros(1, 4).cuda(), torch.zeros(1, 3).cuda())  # camera not set


def test_xfm_ops_match_oracle_and_reference_plugin():
    import diffdope as dd
    from oracle import build_ref, nvdr

    rng = np.random.default_rng(0)
    for Bp, B, N in ((3, 3, 1000), (1, 4, 8240), (2, 2, 1)):
        pts = torch.from_numpy(rng.normal(size=(Bp, N, 3)).astype(np.float32)).cuda().requires_grad_(True)
        M = torch.from_numpy(rng.normal(size=(B, 4, 4)).astype(np.float32)).cuda().requires_grad_(True)
        w = torch.from_numpy(rng.normal(size=(B, N, 4)).astype(np.float32)).cuda()
        out = dd.xfm_points(pts, M)
        assert np.array_equal(out.detach().cpu().numpy(), nvdr.canonical_xfm_points(pts.detach().cpu().numpy(), M.detach().cpu().numpy()))
        (out * w).sum().backward()
        p2, m2 = pts.detach().clone().requires_grad_(True), M.detach().clone().requires_grad_(True)
        (dd.xfm_points(p2, m2, use_python=True) * w).sum().backward()
        assert torch.allclose(pts.grad, p2.grad, rtol=1e-4, atol=1e-4) and torch.allclose(M.grad, m2.grad, rtol=1e-4, atol=1e-2)
        v = dd.xfm_vectors(pts.detach(), M.detach())
        assert torch.allclose(v, dd.xfm_vectors(pts.detach(), M.detach(), use_python=True), rtol=1e-5, atol=1e-5)
    plugin = build_ref.load()  # the reference's own c_src, compiled for sm_100a (oracle/_ref), if it was shipped
    if plugin is not None:
        pts = torch.from_numpy(rng.normal(size=(2, 500, 3)).astype(np.float32)).cuda()
        M = torch.from_numpy(rng.normal(size=(2, 4, 4)).astype(np.float32)).cuda()
        ref = plugin.xfm_fwd(pts, M, True, False)
        assert torch.allclose(dd.xfm_points(pts, M), ref, rtol=1e-6, atol=1e-6)
        gout = torch.from_numpy(rng.normal(size=(2, 500, 4)).astype(np.float32)).cuda()
        ref_dm = plugin.xfm_bwd_mtx(pts, M, gout, True)
        m3 = M.clone().requires_grad_(True)
        (dd.xfm_points(pts, m3) * gout).sum().backward()
        assert torch.allclose(m3.grad, ref_dm, rtol=1e-4, atol=1e-3)


def test_full_size_properties():
    """BASELINE configs[1] size (64 hypotheses, 640^2 window of the 1080p frame): properties that do
    not need the oracle -- bitwise run-to-run determinism, shard equivalence (2 x 32 == 64 with the
    global divisor), loss decrease, finite outputs."""
    ex = Example(1.0, seg1=True)
    win = su.centred_window(ex.gt["segmentation"], 640, ex.H, ex.W)
    ex.sc.set_window(*win)
    B, iters = 64, 25
    lr = torch.from_numpy(su.lr_multipliers(B)).cuda()
    sched = [20.0 * 0.1 ** (it / (iters - 1) + 1) for it in range(iters)]
    cfg = _cfg(ex.n, ALL)

    def run(lo, hi):
        q = torch.from_numpy(np.tile(ex.q, (hi - lo, 1))).cuda().contiguous()
        t = torch.from_numpy(np.tile(ex.t, (hi - lo, 1))).cuda().contiguous()
        ph, lh = ex.sc.optimize(q, t, lr[lo:hi].contiguous(), sched, cfg, b_global=B)
        return torch.cat([q, t], 1), ph, lh

    f1, p1, l1 = run(0, B)
    f2, p2, l2 = run(0, B)
    assert torch.equal(f1, f2) and torch.equal(l1, l2), "fixed reduction order: bitwise reproducible"
    fa, _, la = run(0, 32)
    fb, _, lb = run(32, 64)
    assert torch.equal(torch.cat([fa, fb]), f1) and torch.equal(torch.cat([la, lb], 1), l1), "sharding must not change any hypothesis"
    assert torch.isfinite(f1).all() and torch.isfinite(l1).all()
    total = l1.sum(-1)
    assert (total[-1] < total[0]).float().mean().item() > 0.9, "the loss must go down for almost every hypothesis"
    # prologue + (raster, pixel, iter) per iteration, for each of the two parts a batch of 32 hypotheses is split
    # into (internal streams, include/ddope_b200.h); the split changes no bit of the result (asserted above: 32+32 == 64,
    # where the 64 ran as three parts)
    assert ex.sc.last_launch_count() == 2 * (3 * iters + 1)


def test_backface_culling_rule(ex_half):
    """The library detects the closed example mesh (uv-seam duplicates welded), culls back faces by default, and
    the result equals the no-culling render: identical coverage, identical winners in these views."""
    ex = ex_half
    assert ex.sc.mesh_orientation() == 1 and ex.oracle_mesh().cull_sign == 1
    B = 4
    qs, ts = su.perturbed_poses(ex.q, ex.t, B, rot_deg=25.0)
    qd, td = torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda()
    on = ex.sc.render(qd, td, want=("rast", "rgb", "mask"))
    try:
        ex.sc.set_culling(False)
        off = ex.sc.render(qd, td, want=("rast", "rgb", "mask"))
    finally:
        ex.sc.set_culling(True)
    assert torch.equal(on["rast"][..., 3] > 0, off["rast"][..., 3] > 0)
    assert torch.equal(on["mask"], off["mask"])
    assert float((on["rast"][..., 3] != off["rast"][..., 3]).float().mean()) < 1e-4
    # open mesh (one triangle removed): nothing is culled, both orientations rasterise
    n = ex.n
    sc2 = n.NativeScene(ex.arr["pos"], ex.arr["tri"][:-1], ex.arr["uv"], ex.arr["tex"])
    assert sc2.mesh_orientation() == 0
    # a mirrored model matrix: the other orientation becomes the front, coverage still equals the unculled render
    mtx = on_m = ex.sc.render(qd[:1], td[:1], want=("mtx",))["mtx"].clone()
    mtx[:, :3, 0] *= -1.0
    a = ex.sc.render_mtx(mtx)
    try:
        ex.sc.set_culling(False)
        b = ex.sc.render_mtx(mtx)
    finally:
        ex.sc.set_culling(True)
    assert (a[3][..., 3] > 0).sum() > 1000 and torch.equal(a[3][..., 3] > 0, b[3][..., 3] > 0)
    # camera inside the object's bounding box: back faces are what it sees, nothing may be culled (bit-exact vs the oracle)
    from oracle import refpath

    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32) * 2.0
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    col = np.random.default_rng(3).random((8, 3)).astype(np.float32)
    P = refpath.projection_matrix(60.0, 60.0, 48.0, 32.0, 96, 64)
    cube = n.NativeScene(v, f, vtx_color=col)
    assert cube.mesh_orientation() == 1
    cube.set_camera(P, 64, 96)
    q2 = np.array([[0.1, 0.2, 0.05, 0.97], [0.1, 0.2, 0.05, 0.97]], dtype=np.float32)
    t2 = np.array([[0.1, 0.0, -0.3], [0.0, 0.0, -9.0]], dtype=np.float32)
    og = cube.render(torch.from_numpy(q2).cuda(), torch.from_numpy(t2).cuda(), want=("rast", "rgb"))
    oo = refpath.render(refpath.Mesh(v, f, vtx_color=col), P, torch.from_numpy(q2), torch.from_numpy(t2), 64, 96)
    assert float((og["rast"][0, ..., 3] > 0).float().mean()) > 0.5
    assert np.array_equal(oo["rast_out"].numpy(), og["rast"].cpu().numpy()) and np.array_equal(oo["rgb"].numpy(), og["rgb"].cpu().numpy())


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_triangle_soups_match_oracle_bit_exact(seed):
    """Random open triangle soups (nothing is culled): slivers, strips taller than wide and wider than tall (both
    scanline orientations), triangles larger than 64 px (CTA path), degenerate and duplicated triangles, vertices
    behind the camera and far outside the frame -- coverage, winners, barycentrics, z/w, colour, depth and the
    antialiased mask against the oracle."""
    from oracle import refpath

    n = _nat()
    rng = np.random.default_rng(seed)
    H, W = 80, 112
    P = refpath.projection_matrix(120.0, 115.0, 55.0, 41.0, W, H)
    nt = 260
    ctr = rng.uniform(-1.6, 1.6, (nt, 1, 3)) * np.array([1.3, 1.0, 0.6])
    size = np.exp(rng.uniform(np.log(0.01), np.log(0.9), (nt, 1, 1)))
    v = ctr + rng.normal(0, 1, (nt, 3, 3)) * size
    v[:40, :, 0] = ctr[:40, :, 0] + rng.normal(0, 0.004, (40, 3))           # tall, thin strips
    v[40:80, :, 1] = ctr[40:80, :, 1] + rng.normal(0, 0.004, (40, 3))       # wide, thin strips
    v[80:84] = v[80:81]                                                      # duplicated triangles (lower index wins ties)
    v[84, 2] = v[84, 1]                                                      # degenerate
    v[85:90, :, 2] += 6.0                                                    # behind the camera (w <= 0) -> culled whole
    v[90:93] *= 40.0                                                         # far outside the frame / huge
    pos = v.reshape(-1, 3).astype(np.float32)
    tri = np.arange(nt * 3, dtype=np.int32).reshape(nt, 3)
    col = rng.random((nt * 3, 3)).astype(np.float32)
    sc = n.NativeScene(pos, tri, vtx_color=col)
    assert sc.mesh_orientation() == 0
    sc.set_camera(P, H, W)
    B = 3
    q = rng.normal(0, 1, (B, 4)).astype(np.float32)
    t = (np.array([[0.0, 0.0, -4.5]]) + rng.normal(0, 0.3, (B, 3))).astype(np.float32)
    out = sc.render(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda())
    mesh = refpath.Mesh(pos, tri, vtx_color=col)
    r = refpath.render(mesh, P, torch.from_numpy(q), torch.from_numpy(t), H, W)
    ro, rg = r["rast_out"].numpy(), out["rast"].cpu().numpy()
    assert (ro[..., 3] > 0).mean() > 0.2, "the soup must cover a good part of the frame"
    assert np.array_equal(ro[..., 3], rg[..., 3])
    assert np.array_equal(ro[..., :3], rg[..., :3])
    assert np.array_equal(r["rgb"].numpy(), out["rgb"].cpu().numpy())
    assert np.array_equal(r["depth"].numpy(), out["depth"].cpu().numpy())
    assert np.abs(r["mask"].numpy()[..., 0] - out["mask"].cpu().numpy()).max() <= 1.2e-7
    # and the gradient of the three losses on this soup
    gt = dict(rgb=rng.random((H, W, 3)).astype(np.float32), depth=(4 + rng.random((H, W))).astype(np.float32),
              segmentation=np.repeat((rng.random((H, W, 1)) > 0.4).astype(np.float32), 3, -1))
    g = {k: torch.from_numpy(x).cuda() for k, x in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    lr = su.lr_multipliers(B)
    loss, grad = sc.loss_grad(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(lr).cuda(), _cfg(n, ALL))
    logged, gq, gtr, _ = refpath.forward_backward(mesh, P, q, t, {k: torch.from_numpy(x) for k, x in gt.items()}, lr, ALL, H, W)
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    su.record("triangle_soup_seed%d" % seed, grad_rel_err=su.grad_rel_err(go, gg))
    assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
