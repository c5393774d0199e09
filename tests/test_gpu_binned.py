"""The binned rasterisation path (bin_kernel + TMA-staged per-tile bins + shared-memory z-buffer inside the pixel pass) against the
global z-buffer path: same raster rule, so losses, gradients and whole optimisation runs must be BIT-identical, on micro-triangle
meshes, on triangles larger than a tile, with the edge loss, and when every bin overflows into the scan-the-mesh fallback."""
import numpy as np
import pytest
import torch

import scene_util as su
import workloads as wl
from test_gpu_configs import _gpu_targets, _scene
from test_gpu_parity import ALL, _cfg, _cube_scene, _nat

pytestmark = pytest.mark.gpu

FULL = dict(ALL, l1_edge=True, weight_edge=0.5)


def _both_modes(sc, fn, caps=(None,)):
    sc.set_raster_mode("zbuffer")
    ref = fn()
    outs = []
    for cap in caps:
        sc.set_raster_mode("binned", bin_capacity=cap)
        outs.append(fn())
    sc.set_raster_mode("zbuffer")
    return ref, outs


def _assert_same(ref, outs):
    for got in outs:
        for a, b in zip(ref, got):
            assert torch.equal(a, b)


def test_binned_equals_zbuffer_on_the_bench_geometry():
    """Config 2 geometry (1080p, 640^2 window, 24k covered pixels, ~600 triangles per tile), loss + gradient and a 6-iteration run."""
    n = _nat()
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = su.example_targets(1.0)
    H, W = gt["rgb"].shape[:2]
    sc = n.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(su.projection(), H, W)
    sc.set_window(*su.centred_window(gt["segmentation"], 640, H, W))
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    B = 9  # splits into two parts on internal streams
    qs, ts = su.perturbed_poses(q, t, B, seed=2, rot_deg=2.0, trans=0.02)
    lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 2.0)).cuda()
    for losses in (ALL, FULL, dict(l1_mask=True, weight_mask=1.0)):
        cfg = _cfg(n, losses)

        def one():
            return sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), lr, cfg)

        def run():
            qd, td = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
            ph, lh = sc.optimize(qd, td, lr, [2.0, 1.5, 1.0, 0.8, 0.6, 0.4], cfg)
            return ph, lh, qd, td

        for fn in (one, run):
            ref, outs = _both_modes(sc, fn)
            _assert_same(ref, outs)
    assert sc.bin_overflows() == 0, "the default bin capacity holds every tile of the benchmark geometry"
    # every bin overflows: the tile CTAs scan the whole mesh, results unchanged
    ref, outs = _both_modes(sc, one, caps=(4, 64))
    _assert_same(ref, outs)
    assert sc.bin_overflows() > 0


@pytest.mark.parametrize("big", [False, True])
def test_binned_equals_zbuffer_on_large_triangles(big):
    """A cube: 12 triangles, each spanning many tiles (the 64-bit edge-function path), vertex colours."""
    n, sc, mesh, P, gt, H, W = _cube_scene(H=192, W=256, big=big)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    qs = np.array([[0.3, 0.2, 0.1, 0.9], [0.0, 0.7, 0.1, 0.6]], dtype=np.float32)
    ts = np.array([[0.1, -0.05, -4.0], [-0.3, 0.2, -2.2 if big else -3.0]], dtype=np.float32)
    lr = np.array([1.0, 3.0], dtype=np.float32)
    cfg = _cfg(n, ALL)

    def one():
        return sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), cfg)

    ref, outs = _both_modes(sc, one, caps=(None, 4))
    _assert_same(ref, outs)


def test_binned_equals_zbuffer_on_the_stress_workload():
    """Config 5 stand-in at full size (50k triangles, half of a 1024^2 window covered), 4 hypotheses, full stack incl. edge loss."""
    n = _nat()
    w = wl.config5()
    sc = _scene(n, w)
    g = _gpu_targets(sc, w)
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    B = 4
    q = torch.from_numpy(np.tile(w["q0"], (B, 1))).cuda().contiguous()
    t = torch.from_numpy(np.tile(w["t0"], (B, 1))).cuda().contiguous()
    lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 2.0)).cuda()
    cfg = _cfg(n, w["losses"])

    def run():
        qd, td = q.clone(), t.clone()
        ph, lh = sc.optimize(qd, td, lr, [2.0, 1.0, 0.5], cfg)
        return ph, lh, qd, td

    ref, outs = _both_modes(sc, run)
    _assert_same(ref, outs)
    assert sc.bin_overflows() == 0


def test_small_batch_graph_replay_is_bit_identical():
    """Fewer than 8 hypotheses: ddope_optimize replays its launches as one CUDA graph (updated in place from call to call, rebuilt when
    the iteration count changes); same tables as the direct launches, in both raster modes, from the legacy default stream and from
    a side stream."""
    n = _nat()
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = su.example_targets(0.5)
    H, W = gt["rgb"].shape[:2]
    sc = n.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(su.projection(), H, W)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    cfg = _cfg(n, ALL)
    for mode in ("zbuffer", "binned"):
        sc.set_raster_mode(mode)
        for B, iters in ((1, 12), (3, 12), (3, 7), (1, 12)):
            qs, ts = su.perturbed_poses(q, t, B, seed=B, rot_deg=2.0, trans=0.02)
            lr = torch.from_numpy(su.lr_multipliers(B, 0.01, 1.0)).cuda()
            sched = [2.0 * 0.9 ** i for i in range(iters)]

            def run():
                qd, td = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
                ph, lh = sc.optimize(qd, td, lr, sched, cfg)
                torch.cuda.synchronize()
                return ph, lh, qd, td

            sc.set_graph(False)
            ref = run()
            sc.set_graph(True)
            before = sc.graph_launch_count()
            got = run()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                got2 = run()
            torch.cuda.current_stream().wait_stream(side)
            assert sc.graph_launch_count() == before + 2, "both calls were served by a graph launch"
            for a, b, c in zip(ref, got, got2):
                assert torch.equal(a, b) and torch.equal(a, c)
    sc.set_raster_mode("zbuffer")
