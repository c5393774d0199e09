"""CPU: pin the oracle (oracle/) against every number that exists for this path.

The reference has no tests and no golden vectors; its hot-path arithmetic
lives in nvdiffrast, which is not available, so parity is UNPINNED for those ops. What can be
pinned is pinned here: SURVEY.md Appendix D (numbers derived from the reference's own data and
loaders), the reference's `use_python` xfm formula, finite-difference checks of the hand-written
backward formulas, and the committed golden fixture (regression)."""
import os
import random

import numpy as np
import pytest
import torch

import scene_util as su
from oracle import nvdr, refpath

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "example_q25.npz")


def test_pin_legacy_rotation_block_is_identity():
    # Appendix D.1: Rz(pi/2) Ry(-pi/2) Rz(-pi/2) Rx(-pi/2) = I  (diffdope.py:128-137)
    from diffdope._quat import quat_axis, quat_mul, quat_to_matrix33

    q = quat_mul(quat_mul(quat_mul(quat_axis("z", np.pi / 2), quat_axis("y", -np.pi / 2)), quat_axis("z", -np.pi / 2)), quat_axis("x", -np.pi / 2))
    assert np.allclose(quat_to_matrix33(q), np.eye(3), atol=1e-12)


def test_pin_config_pose():
    # Appendix D.2
    q, t = su.example_pose()
    assert np.allclose(t, [-1.6116878, -2.0622094, -7.47151334], atol=1e-6)
    ref = np.array([0.28427788, -0.34248786, 0.88225564, -0.15333994])
    assert np.allclose(q, ref, atol=1e-6) or np.allclose(q, -ref, atol=1e-6)


def test_pin_projection():
    # Appendix D.3 (diffdope.py:679-742, y_down branch)
    P = su.projection()
    ref = np.array([[1.44846875, 0, -0.00516354, 0], [0, 2.5685, -0.03224815, 0], [0, 0, -1.00010001, -0.020001], [0, 0, -1, 0]])
    assert np.allclose(P, ref, atol=1e-7)
    assert np.allclose(P, su.projection_native(), atol=0)


def test_pin_mesh():
    # Appendix D.4
    a = su.example_mesh_arrays()
    assert a["pos"].shape == (8240, 3) and a["tri"].shape == (13860, 3) and a["tri"].max() == 8239
    assert np.allclose(np.abs(a["pos"]).max(0), [0.355603, 0.330279, 0.417773], atol=1e-6)
    v = 1 - a["uv"][:, 1]
    assert abs(a["uv"][:, 0].min() - 0.002) < 1e-4 and abs(a["uv"][:, 0].max() - 0.998) < 1e-4
    assert abs(v.min() - 0.0018) < 1e-4 and abs(v.max() - 0.9117) < 1e-4


def test_pin_geometry_at_config_pose():
    # Appendix D.5: clip w range, NDC z range, screen bbox at 960x540, facing counts, area sum
    a = su.example_mesh_arrays()
    q, t = su.example_pose()
    _, M = nvdr.canonical_pose(q[None], t[None])
    clip = nvdr.canonical_xfm_points(a["pos"], nvdr.canonical_mvp(su.projection(), M))[0]
    w = clip[:, 3]
    assert abs(w.min() - 6.9743) < 2e-4 and abs(w.max() - 7.9816) < 2e-4
    z = clip[:, 2] / w
    assert abs(z.min() - 0.997232) < 2e-6 and abs(z.max() - 0.997594) < 2e-6
    sx = (clip[:, 0] / w * 0.5 + 0.5) * 960
    sy = (clip[:, 1] / w * 0.5 + 0.5) * 540
    assert abs(sx.min() - 290.52) < 0.02 and abs(sx.max() - 381.09) < 0.02
    assert abs(sy.min() - 36.57) < 0.02 and abs(sy.max() - 136.64) < 0.02
    tri = a["tri"]
    area = 0.5 * ((sx[tri[:, 1]] - sx[tri[:, 0]]) * (sy[tri[:, 2]] - sy[tri[:, 0]]) - (sx[tri[:, 2]] - sx[tri[:, 0]]) * (sy[tri[:, 1]] - sy[tri[:, 0]]))
    assert (area > 0).sum() == 7081 and (area < 0).sum() == 6779
    assert abs(area[area > 0].sum() - 6248.94) < 0.5
    assert abs(np.median(np.abs(area)) - 0.4197) < 2e-3


def test_pin_targets():
    # Appendix D.6 / D.7: image pipeline of diffdope.py:1122-1152 at 0.5x
    gt = su.example_targets(0.5)
    seg = gt["segmentation"]
    assert seg.shape == (540, 960, 3)
    assert np.array_equal(seg[..., 0], seg[..., 1]) and np.array_equal(seg[..., 0], seg[..., 2])
    assert set(np.unique(seg)) <= {0.0, 0.25, 0.5, 0.75, 1.0}
    assert abs(seg[..., 0].sum() - 5440.25) < 1e-3
    ys, xs = np.nonzero(seg[..., 0])
    assert (ys.min(), ys.max(), xs.min(), xs.max()) == (44, 140, 295, 383)
    d = gt["depth"]
    inside = d[seg[..., 0] > 0.5]
    assert inside.min() == 0.0 and abs(np.median(inside) - 7.49) < 0.01 and abs(inside.max() - 8.38) < 0.01
    assert abs((inside == 0).mean() - 0.0153) < 1e-3


def test_pin_schedule_and_multipliers():
    # Appendix D.8 (diffdope.py:1657-1661, 1368-1374)
    lrs = [refpath.lr_schedule(it, 60, 20, 0.1) for it in range(61)]
    assert len(lrs) == 61 and abs(lrs[0] - 2.0) < 1e-12 and abs(lrs[-1] - 0.2) < 1e-12
    m = su.lr_multipliers(3)
    assert np.allclose(m, [84.44374093398956, 75.79786075000085, 42.06295236727619], rtol=1e-7)


def test_pin_abs_subgradient():
    # Appendix D.9
    x = torch.zeros(3, requires_grad=True)
    torch.abs(x).sum().backward()
    assert torch.all(x.grad == 0)


def test_xfm_matches_reference_python_path():
    # diffdope/ops.py:137-141 (the reference's own torch validation formula)
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(3, 50, 3)).astype(np.float32)
    M = rng.normal(size=(3, 4, 4)).astype(np.float32)
    ref = torch.matmul(torch.nn.functional.pad(torch.from_numpy(pts), (0, 1), value=1.0), torch.from_numpy(M).transpose(1, 2)).numpy()
    assert np.allclose(nvdr.canonical_xfm_points(pts, M), ref, rtol=1e-5, atol=1e-5)
    p = torch.from_numpy(pts).requires_grad_(True)
    m = torch.from_numpy(M).requires_grad_(True)
    w = torch.from_numpy(rng.normal(size=(3, 50, 4)).astype(np.float32))
    (refpath.xfm_points(p, m) * w).sum().backward()
    p2 = torch.from_numpy(pts).requires_grad_(True)
    m2 = torch.from_numpy(M).requires_grad_(True)
    (torch.matmul(torch.nn.functional.pad(p2, (0, 1), value=1.0), m2.transpose(1, 2)) * w).sum().backward()
    assert np.allclose(p.grad, p2.grad, rtol=1e-4, atol=1e-5) and np.allclose(m.grad, m2.grad, rtol=1e-4, atol=1e-4)


# ----------------------------------------------------------------------------------------------
# raster rule properties


def _quad_mesh():
    v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    return v, np.array([[0, 1, 2], [0, 2, 3]])


def test_raster_shared_edge_watertight_and_exclusive():
    """Every pixel centre inside a quad split along its diagonal is covered exactly once whichever
    way the triangles wind (tie rule), including centres exactly on the diagonal."""
    H = W = 16
    for flip in (False, True):
        clip = np.array([[[-0.5, -0.5, 0, 1], [0.5, -0.5, 0, 1], [0.5, 0.5, 0, 1], [-0.5, 0.5, 0, 1]]], dtype=np.float32)
        tri = np.array([[0, 1, 2], [0, 2, 3]])
        if flip:
            tri = tri[:, ::-1].copy()
        # count coverage per triangle separately
        cov = np.zeros((H, W), int)
        for k in range(2):
            r = nvdr.rasterize(clip, tri[k:k + 1], H, W)
            cov += (r[0, ..., 3] > 0)
        inside = cov[4:12, 4:12]
        assert inside.min() == 1 and inside.max() == 1, "diagonal pixels must belong to exactly one triangle"
        assert cov.sum() == 64


def test_raster_depth_less_and_first_wins():
    H = W = 8
    clip = np.array([[[-1, -1, 0.5, 1], [1, -1, 0.5, 1], [0, 1, 0.5, 1], [-1, -1, 0.2, 1], [1, -1, 0.2, 1], [0, 1, 0.2, 1]]], dtype=np.float32)
    tri = np.array([[0, 1, 2], [3, 4, 5], [3, 4, 5]])
    r = nvdr.rasterize(clip, tri, H, W)
    ids = r[0, ..., 3]
    assert set(np.unique(ids)) == {0.0, 2.0}, "nearer triangle wins; equal depth keeps the lower index"
    assert np.allclose(r[0, ..., 2][ids > 0], 0.2)


def test_raster_culls_behind_camera_and_zclip():
    H = W = 8
    clip = np.array([[[-1, -1, 0, 1], [1, -1, 0, 1], [0, 1, 0, -1]]], dtype=np.float32)
    assert nvdr.rasterize(clip, np.array([[0, 1, 2]]), H, W)[..., 3].max() == 0
    clip = np.array([[[-1, -1, 2, 1], [1, -1, 2, 1], [0, 1, 2, 1]]], dtype=np.float32)
    assert nvdr.rasterize(clip, np.array([[0, 1, 2]]), H, W)[..., 3].max() == 0


def test_barycentrics_reconstruct_attributes():
    H = W = 32
    clip = np.array([[[-0.8, -0.7, 0.1, 1.0], [0.9, -0.6, 0.3, 2.0], [0.1, 0.8, 0.2, 1.5]]], dtype=np.float32)
    tri = np.array([[0, 1, 2]])
    r = nvdr.rasterize(clip, tri, H, W)
    cov = r[0, ..., 3] > 0
    assert cov.sum() > 100
    # interpolating clip x/w... : interpolate w-weighted positions must reproduce pixel NDC
    attr = clip[0, :, :4]
    out = nvdr.interpolate(attr, r, tri)[0]
    py, px = np.nonzero(cov)
    fx, fy = nvdr.pixel_ndc(px, py, W, H)
    assert np.allclose(out[py, px, 0] / out[py, px, 3], fx, atol=2e-5)
    assert np.allclose(out[py, px, 1] / out[py, px, 3], fy, atol=2e-5)
    assert np.allclose(out[py, px, 2] / out[py, px, 3], r[0, py, px, 2], atol=2e-5)


def test_texture_linear_texel_centres_and_wrap():
    tex = np.arange(4 * 4 * 3, dtype=np.float32).reshape(4, 4, 3)
    uv = np.array([[[[(2 + 0.5) / 4, (1 + 0.5) / 4]]]], dtype=np.float32)
    assert np.allclose(nvdr.texture_linear(tex, uv)[0, 0, 0], tex[1, 2])
    uv = np.array([[[[0.0, 0.0]]]], dtype=np.float32)  # corner: average of the four wrapped corner texels
    assert np.allclose(nvdr.texture_linear(tex, uv)[0, 0, 0], (tex[0, 0] + tex[0, 3] + tex[3, 0] + tex[3, 3]) / 4)


def test_edge_opposites():
    tri = np.array([[0, 1, 2], [0, 2, 3]])
    opp = nvdr.build_edge_opposites(tri)
    # triangle 0: edge opposite v1 (index 1) is (2,0), shared with triangle 1 whose other vertex is 3
    assert opp[0].tolist() == [-1, 3, -1]
    assert opp[1].tolist() == [-1, -1, 1]


# ----------------------------------------------------------------------------------------------
# gradient checks (finite differences through the full restated graph)


def _cube():
    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32) * 0.5
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
    col = np.random.default_rng(3).random((8, 3)).astype(np.float32)
    return v, f, col


@pytest.mark.parametrize("which", ["mask", "depth", "rgb"])
def test_analytic_gradient_matches_finite_differences(which):
    v, f, col = _cube()
    mesh = refpath.Mesh(v, f, vtx_color=col)
    H, W = 64, 96
    P = refpath.projection_matrix(100.0, 100.0, 48.0, 32.0, 96, 64)
    yy, xx = np.mgrid[0:H, 0:W]
    wgt = torch.tensor((np.sin(xx / 7.0) + np.cos(yy / 5.0)).astype(np.float32))
    q0 = np.array([[0.3, 0.2, 0.1, 0.9]], dtype=np.float32)
    t0 = np.array([[0.1, -0.05, -4.0]], dtype=np.float32)

    def L(q, t, grad=False):
        qq = torch.tensor(q, requires_grad=True)
        tt = torch.tensor(t, requires_grad=True)
        r = refpath.render(mesh, P, qq, tt, H, W)
        if which == "mask":
            l = (r["mask"][0, ..., 0] * wgt).sum()
        elif which == "depth":
            l = (r["depth"][0] * wgt * (r["rast_out"][0, ..., 3] > 0)).sum() * 0 + ((r["depth"][0] + 4.0) * wgt * (r["rast_out"][0, ..., 3].detach() > 0)).sum()
        else:
            l = (r["rgb"][0].sum(-1) * wgt).sum()
        if grad:
            l.backward()
            return float(l), qq.grad.numpy().copy(), tt.grad.numpy().copy()
        return float(l)

    _, gq, gt = L(q0, t0, True)
    h = 2e-3
    # z translation and the quaternion w component move the silhouette least: compare where finite
    # differences are stable (interior-dominated for depth/rgb, AA-resolved for mask)
    fd_t = []
    for i in range(3):
        tp, tm = t0.copy(), t0.copy()
        tp[0, i] += h
        tm[0, i] -= h
        fd_t.append((L(q0, tp) - L(q0, tm)) / (2 * h))
    fd_t = np.array(fd_t)
    if which == "mask":
        # antialiased coverage is piecewise linear in the silhouette position: FD and analytic agree
        assert np.allclose(gt[0], fd_t, rtol=0.25, atol=0.08 * np.abs(fd_t).max())
    else:
        # no antialias on rgb/depth (as in the reference): silhouette popping makes FD noisy in x,y;
        # the depth direction barely moves the silhouette
        assert abs(gt[0, 2] - fd_t[2]) < 0.35 * max(abs(fd_t[2]), 1e-3) + 0.05 * np.abs(fd_t).max()


# ----------------------------------------------------------------------------------------------
# golden fixture (regression pin of the oracle itself)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_oracle_reproduces_golden(golden):
    g = golden
    arr = su.example_mesh_arrays()
    gt = {k: torch.from_numpy(v) for k, v in su.example_targets(float(g["resize"])).items()}
    H, W = int(g["H"]), int(g["W"])
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    cfg = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
    logged, gq, gtr, r = refpath.forward_backward(mesh, su.projection(), g["quat"], g["trans"], gt, g["lr"], cfg, H, W)
    assert np.array_equal(r["rast_out"].detach().numpy()[..., 3].astype(np.int32), g["tri_id"])
    y0, y1, x0, x1 = g["bbox"]
    assert np.array_equal(r["rgb"].detach().numpy()[:, y0:y1, x0:x1], g["rgb"])
    assert np.array_equal(r["depth"].detach().numpy()[:, y0:y1, x0:x1], g["depth"])
    loss = np.stack([logged["rgb"].numpy(), logged["depth"].numpy(), logged["mask_selection"].numpy()], 1)
    assert np.allclose(loss, g["loss"], rtol=1e-6)
    assert np.allclose(np.concatenate([gq, gtr], 1), g["grad"], rtol=1e-4, atol=1e-7)


def test_window_equals_slice_of_full_frame(golden):
    """Loss over a window = loss over the slice of the full-frame render."""
    g = golden
    arr = su.example_mesh_arrays()
    gt = {k: torch.from_numpy(v) for k, v in su.example_targets(float(g["resize"])).items()}
    H, W = int(g["H"]), int(g["W"])
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    cfg = dict(l1_mask=True, weight_mask=1.0)
    win = su.centred_window(gt["segmentation"].numpy(), 96, H, W)
    q = torch.tensor(g["quat"][:1])
    t = torch.tensor(g["trans"][:1])
    r = refpath.render(mesh, su.projection(), q, t, H, W)
    _, logged = refpath.losses(r, gt, torch.ones(1), cfg, window=win)
    y0, x0, h, w = win
    manual = torch.abs(r["mask"][:, y0:y0 + h, x0:x0 + w] - gt["segmentation"][None, y0:y0 + h, x0:x0 + w]).mean((1, 2, 3))
    assert torch.allclose(logged["mask_selection"], manual)


# ----------------------------------------------------------------------------------------------
# extensions (no reference counterpart): pinned against closed forms


def test_ext_sobel_closed_forms():
    """Linear ramps have constant Sobel response 8*slope in the interior; the window border is zero-padded."""
    H, W = 12, 16
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = torch.tensor(np.stack([0.5 * xx + 0.25 * yy] * 3, -1)[None])
    e = refpath.sobel_magnitude(img)[0].numpy()
    assert np.allclose(e[1:-1, 1:-1], np.hypot(8 * 0.5, 8 * 0.25), rtol=1e-6)
    assert e[0, 0] != pytest.approx(np.hypot(4.0, 2.0))  # border sees the zero padding
    flat = refpath.sobel_magnitude(torch.full((1, H, W, 3), 0.7))[0].numpy()
    assert np.allclose(flat[1:-1, 1:-1], 1e-6, atol=1e-8)  # sqrt(1e-12)
    # grey is the channel mean
    rgb = torch.rand(1, H, W, 3)
    assert torch.allclose(refpath.sobel_magnitude(rgb), refpath.sobel_magnitude(rgb.mean(-1, keepdim=True).expand(-1, -1, -1, 3)), atol=1e-6)


def test_ext_mip_chain_and_lod_closed_form():
    tex = np.random.default_rng(0).random((64, 32, 3)).astype(np.float32)
    lv = nvdr.build_mip_chain(tex)
    assert [l.shape[:2] for l in lv] == [(64, 32), (32, 16), (16, 8), (8, 4), (4, 2), (2, 1), (1, 1)]
    for l in lv:
        assert np.allclose(l.mean((0, 1)), tex.mean((0, 1)), atol=1e-6)  # box filtering preserves the mean
    assert np.allclose(lv[1][3, 5], tex[6:8, 10:12].mean((0, 1)), atol=1e-7)
    assert len(nvdr.build_mip_chain(tex, max_levels=3)) == 3
    # fronto-parallel unit quad at depth d, uv = (x+0.5, y+0.5): footprint = tex_size / pixels-across
    fx = 100.0
    H = W = 64
    P = refpath.projection_matrix(fx, fx, 32.0, 32.0, W, H)
    d = 4.0
    pos = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0.5, 0.5, 0], [-0.5, 0.5, 0]], dtype=np.float32)
    tri = np.array([[0, 1, 2], [0, 2, 3]])
    uv = (pos[:, :2] + 0.5).astype(np.float32)
    M = np.eye(4, dtype=np.float32)[None].copy()
    M[0, 2, 3] = -d
    clip = nvdr.canonical_xfm_points(pos[None], nvdr.canonical_mvp(P, M))
    rast = nvdr.rasterize(clip, tri, H, W)
    assert (rast[..., 3] > 0).sum() > 400
    for tw in (256, 1024):
        lod = nvdr.texture_lod(clip, tri, uv, rast, (tw, tw), 11)
        want = np.log2(tw / (fx / d))  # texels per pixel: tw texels over fx/d pixels
        cov = rast[..., 3] > 0
        assert np.allclose(lod[cov], want, atol=1e-3)
        assert (lod[~cov] == 0).all()
    # trilinear lookup of a constant-per-level chain returns the level blend
    levels = [np.full((8 >> l, 8 >> l, 3), float(l), np.float32) for l in range(4)]
    uvq = np.random.default_rng(1).random((1, 4, 4, 2)).astype(np.float32)
    lodq = np.array([0.0, 0.25, 1.5, 3.0, 2.999, 0.5, 1.0, 2.0] * 2, np.float32).reshape(1, 4, 4)
    out = nvdr.texture_mipmap(levels, uvq, lodq)
    assert np.allclose(out[..., 0], lodq, atol=1e-6)


def test_ext_mipmap_uv_gradient_matches_finite_differences():
    rng = np.random.default_rng(2)
    levels = nvdr.build_mip_chain(rng.random((32, 32, 3)).astype(np.float32))
    uv = rng.random((1, 6, 6, 2)).astype(np.float32) * 3 - 1  # wraps
    lod = (rng.random((1, 6, 6)) * 3.5).astype(np.float32)
    dy = rng.normal(size=(1, 6, 6, 3)).astype(np.float32)
    g = nvdr.texture_mipmap_grad_uv(levels, uv, lod, dy)
    h = 1e-3
    for k in range(2):
        up, um = uv.copy(), uv.copy()
        up[..., k] += h
        um[..., k] -= h
        fd = ((nvdr.texture_mipmap(levels, up, lod) - nvdr.texture_mipmap(levels, um, lod)) * dy).sum(-1) / (2 * h)
        # bilinear is piecewise linear in uv: exclude samples whose +-h interval crosses a texel boundary
        ok = np.ones(fd.shape, bool)
        for l in range(len(levels)):
            s = levels[l].shape[1 - k]
            a = np.floor((up[..., k] - np.floor(up[..., k])) * s - 0.5)
            b = np.floor((um[..., k] - np.floor(um[..., k])) * s - 0.5)
            ok &= a == b
        assert ok.sum() > 10
        assert np.allclose(g[..., k][ok], fd[ok], rtol=2e-2, atol=2e-2)


def test_ext_adam_oracle_first_step_is_lr():
    v, f, col = _cube()
    mesh = refpath.Mesh(v, f, vtx_color=col)
    H, W = 48, 64
    P = refpath.projection_matrix(80.0, 80.0, 32.0, 24.0, W, H)
    rng = np.random.default_rng(5)
    gt = dict(rgb=torch.tensor(rng.random((H, W, 3)).astype(np.float32)), depth=torch.tensor((3 + rng.random((H, W))).astype(np.float32)),
              segmentation=torch.ones(H, W, 3))
    q0 = np.array([[0.3, 0.2, 0.1, 0.9]], dtype=np.float32)
    t0 = np.array([[0.1, -0.05, -4.0]], dtype=np.float32)
    losses = dict(l1_rgb_with_mask=True, l1_depth_with_mask=True, l1_mask=True)
    hyper = dict(nb_iterations=1, base_lr=0.01, lr_decay=0.5, learning_rate_base=1, optimizer="adam")
    o = refpath.run_optimization(mesh, P, q0, t0, gt, np.ones(1, np.float32), losses, hyper, H, W)
    lr0 = refpath.lr_schedule(0, 1, 0.01, 0.5)
    assert np.allclose(np.abs(o["poses"][1] - o["poses"][0]), lr0, rtol=1e-3)


def test_closed_mesh_detection_and_culling_preserves_coverage():
    """Raster rule: back faces of a closed, consistently oriented mesh are skipped. Coverage must not change,
    and the winner may differ from the no-culling render only on silhouette ties (none in these views)."""
    v, f, col = _cube()
    assert nvdr.closed_mesh_orientation(v, f) == 1
    assert nvdr.closed_mesh_orientation(v, f[:, ::-1]) == -1          # inside-out
    assert nvdr.closed_mesh_orientation(v, f[:-1]) == 0               # a hole
    g = f.copy()
    g[3] = g[3, ::-1]
    assert nvdr.closed_mesh_orientation(v, g) == 0                    # one flipped triangle
    # duplicated seam vertices (same position, different index) still weld into a closed surface
    v2 = np.concatenate([v, v[:1]])
    f2 = f.copy()
    f2[0, 0] = 8
    assert nvdr.closed_mesh_orientation(v2, f2) == 1
    arr = su.example_mesh_arrays()
    assert nvdr.closed_mesh_orientation(arr["pos"], arr["tri"]) == 1  # 2548 index-level seam edges, closed after welding
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    q, t = su.example_pose()
    qs, ts = su.perturbed_poses(q, t, 4, rot_deg=25.0)
    P = su.projection()
    H, W = 135, 240
    r1 = refpath.render(mesh, P, torch.from_numpy(qs), torch.from_numpy(ts), H, W)["rast_out"].numpy()
    mesh.cull = False
    r0 = refpath.render(mesh, P, torch.from_numpy(qs), torch.from_numpy(ts), H, W)["rast_out"].numpy()
    assert (r0[..., 3] > 0).sum() > 1000
    assert np.array_equal(r0[..., 3] > 0, r1[..., 3] > 0)
    assert (r0[..., 3] != r1[..., 3]).mean() < 1e-4
    # a mirrored model matrix flips which orientation is the front
    M = np.eye(4)[None].repeat(2, 0)
    M[1, 0, 0] = -1
    assert list(nvdr.face_signs(1, P, M)) == [1, -1]
    # a camera inside the object's bounding box sees back faces: nothing is culled for that hypothesis
    big = refpath.Mesh(v * 4, f, vtx_color=col)  # cube of half-size 2
    q2 = np.array([[0.1, 0.2, 0.05, 0.97], [0.1, 0.2, 0.05, 0.97]], dtype=np.float32)
    t2 = np.array([[0.1, 0.0, -0.3], [0.0, 0.0, -9.0]], dtype=np.float32)
    _, M2 = nvdr.canonical_pose(q2, t2)
    assert list(nvdr.face_signs(1, P, M2, big.pos.min(0), big.pos.max(0))) == [0, 1]
    r2 = refpath.render(big, P, torch.from_numpy(q2), torch.from_numpy(t2), 64, 96)["rast_out"].numpy()
    assert (r2[0, ..., 3] > 0).mean() > 0.5 and (r2[1, ..., 3] > 0).mean() > 0.02


# ----------------------------------------------------------------------------------------------
# independent cross-checks of the oracle's geometry (nvdiffrast itself is not available: these pin the
# restated rasterise / interpolate / depth arithmetic against first-principles computations instead)


def _raycast(pos_cam, tri, P, H, W):
    """Brute-force float64 ray casting through every pixel centre (Moeller-Trumbore): front-most triangle id,
    3-D barycentrics (= perspective-correct attribute weights) and camera-space depth -z."""
    P = np.asarray(P, dtype=np.float64)
    xs = (np.arange(W) + 0.5) * 2.0 / W - 1.0
    ys = (np.arange(H) + 0.5) * 2.0 / H - 1.0
    X, Y = np.meshgrid(xs, ys)
    # clip = P [x,y,z,1], w = -z: x_ndc = (P00 x + P02 z)/(-z), y_ndc = (P11 y + P12 z)/(-z); on the plane z = -1
    # that is x_ndc = P00 x - P02, so the ray through (x_ndc, y_ndc) has direction ((x_ndc + P02)/P00, (y_ndc + P12)/P11, -1)
    d = np.stack([(X + P[0, 2]) / P[0, 0], (Y + P[1, 2]) / P[1, 1], -np.ones_like(X)], -1).reshape(-1, 3)
    best_t = np.full(d.shape[0], np.inf)
    best_id = np.full(d.shape[0], -1)
    best_uv = np.zeros((d.shape[0], 2))
    for k, (a, b, c) in enumerate(tri):
        v0, v1, v2 = pos_cam[a], pos_cam[b], pos_cam[c]
        e1, e2 = v1 - v0, v2 - v0
        pv = np.cross(d, e2)
        det = pv @ e1
        with np.errstate(all="ignore"):
            inv = 1.0 / det
            tv = -v0  # ray origin is the camera centre
            bu = (pv @ tv) * inv
            qv = np.cross(tv, e1)
            bv = (d @ qv) * inv
            t = (qv @ e2) * inv
        hit = (np.abs(det) > 1e-12) & (bu >= 0) & (bv >= 0) & (bu + bv <= 1) & (t > 0) & (t < best_t)
        best_t[hit] = t[hit]
        best_id[hit] = k
        # nvdiffrast's (u, v) are the weights of vertex 0 and vertex 1
        best_uv[hit, 0] = (1 - bu - bv)[hit]
        best_uv[hit, 1] = bu[hit]
    depth = np.where(best_id >= 0, best_t, 0.0)  # direction has z = -1: t is exactly -z
    return best_id.reshape(H, W), best_uv.reshape(H, W, 2), depth.reshape(H, W)


def test_oracle_matches_ray_casting():
    v, f, col = _cube()
    mesh = refpath.Mesh(v, f, vtx_color=col)
    H, W = 72, 96
    P = refpath.projection_matrix(110.0, 105.0, 47.3, 37.1, W, H)
    q = np.array([[0.31, -0.22, 0.12, 0.91]], dtype=np.float32)
    t = np.array([[0.12, -0.07, -3.5]], dtype=np.float32)
    r = refpath.render(mesh, P, torch.from_numpy(q), torch.from_numpy(t), H, W)
    M = r["mtx"][0].numpy().astype(np.float64)
    pos_cam = v.astype(np.float64) @ M[:3, :3].T + M[:3, 3]
    ids, uv, depth = _raycast(pos_cam, f, P, H, W)
    rast = r["rast_out"][0].numpy()
    oid = rast[..., 3].astype(np.int64) - 1
    assert (oid >= 0).sum() > 400
    # interior pixels (same triangle in the 3x3 neighbourhood of the ray-cast image) are away from every edge, where
    # neither the 1/256 px snapping nor a tie rule can matter: there the two must agree exactly on the winner
    interior = np.ones_like(ids, bool)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            interior &= np.roll(np.roll(ids, dy, 0), dx, 1) == ids
    interior &= ids >= 0
    assert interior.sum() > 250
    assert np.array_equal(oid[interior], ids[interior])
    assert np.abs(rast[..., 0][interior] - uv[..., 0][interior]).max() < 2e-5
    assert np.abs(rast[..., 1][interior] - uv[..., 1][interior]).max() < 2e-5
    assert np.abs(r["depth"][0].numpy()[interior] - depth[interior]).max() < 2e-5
    # everywhere: coverage differs only on the silhouette ring (snapping), winners only next to an edge
    cov_diff = (oid >= 0) != (ids >= 0)
    assert cov_diff.sum() <= 0.02 * (ids >= 0).sum()
    # interpolated vertex colour = barycentric blend of the ray-cast weights
    want = (uv[..., 0:1] * col[f[ids.clip(0), 0]] + uv[..., 1:2] * col[f[ids.clip(0), 1]] + (1 - uv[..., 0:1] - uv[..., 1:2]) * col[f[ids.clip(0), 2]])
    assert np.abs(r["rgb"][0].numpy()[interior] - want[interior]).max() < 3e-5


def test_triangles_crossing_the_camera_plane_match_ray_casting():
    """Near-plane handling (GL / nvdiffrast clip; the oracle rasterises such triangles in homogeneous coordinates, DESIGN.md
    section 4): a ground plane under the camera whose far end is in view and whose near end lies BEHIND the camera plane (two of
    its four vertices have w < 0), and a slanted triangle with one vertex behind. Winner, barycentrics and depth against float64
    ray casting; nothing appears above the horizon or behind the camera; a triangle entirely behind is dropped."""
    v = np.array([[-2.0, -0.3, 1.0], [2.0, -0.3, 1.0], [2.0, -0.3, -6.0], [-2.0, -0.3, -6.0],
                  [0.3, 0.5, -2.0], [0.9, 0.1, -1.5], [0.5, 0.9, 0.7],
                  [-1.0, 0.2, 0.5], [-0.5, 0.3, 2.0], [-0.8, 0.9, 1.0]], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.int32)
    col = np.random.default_rng(5).random((10, 3)).astype(np.float32)
    mesh = refpath.Mesh(v, f, vtx_color=col)
    assert mesh.cull_sign == 0
    H, W = 72, 96
    P = refpath.projection_matrix(70.0, 75.0, 47.3, 37.1, W, H)
    for q in ([0.0, 0.0, 0.0, 1.0], [0.05, -0.08, 0.03, 0.99]):
        q = np.array([q], dtype=np.float32)
        t = np.array([[0.02, -0.03, 0.0]], dtype=np.float32)
        r = refpath.render(mesh, P, torch.from_numpy(q), torch.from_numpy(t), H, W)
        M = r["mtx"][0].numpy().astype(np.float64)
        pos_cam = v.astype(np.float64) @ M[:3, :3].T + M[:3, 3]
        assert (pos_cam[f[0], 2] > 0).sum() == 2 and (pos_cam[f[2], 2] > 0).sum() == 1 and (pos_cam[f[3], 2] > 0).all()
        ids, uv, depth = _raycast(pos_cam, f, P, H, W)
        rast = r["rast_out"][0].numpy()
        oid = rast[..., 3].astype(np.int64) - 1
        assert (oid == 0).sum() + (oid == 1).sum() > 1000 and (oid == 2).sum() > 50, "the crossing triangles are rendered, not dropped"
        assert not (oid == 3).any(), "a triangle entirely behind the camera plane is dropped"
        interior = np.ones_like(ids, bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                interior &= np.roll(np.roll(ids, dy, 0), dx, 1) == ids
        interior[[0, -1], :] = False
        interior[:, [0, -1]] = False
        cov = interior & (ids >= 0)
        assert cov.sum() > 900
        assert np.array_equal(oid[interior], ids[interior]), "same winner (and same background) away from edges"
        assert np.abs(rast[..., 0][cov] - uv[..., 0][cov]).max() < 1e-4 and np.abs(rast[..., 1][cov] - uv[..., 1][cov]).max() < 1e-4
        assert np.abs(r["depth"][0].numpy()[cov] - depth[cov]).max() < 1e-4 * depth[cov].max()
        assert ((oid >= 0) != (ids >= 0)).sum() <= 0.03 * (ids >= 0).sum()


def test_oracle_coverage_matches_opencv_polygon_fill():
    """Silhouette of the example mesh against OpenCV's polygon rasteriser fed with the same 1/256 px fixed-point
    vertices (an independent implementation with a different fill rule: boundary pixels may differ)."""
    import cv2

    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    H, W = 270, 480
    P = su.projection()
    qh, M = nvdr.canonical_pose(q[None], t[None])
    clip = nvdr.canonical_xfm_points(arr["pos"][None], nvdr.canonical_mvp(P, M))
    cov = nvdr.rasterize(clip, arr["tri"], H, W)[0, ..., 3] > 0
    X, Y, ok = nvdr.snap_vertices(clip[0], W, H)
    assert ok.all()
    img = np.zeros((H, W), np.uint8)
    tri = arr["tri"]
    # pixel centres sit at +0.5: shift by half a pixel so that cv2's integer pixel grid samples centres
    pts = np.stack([X[tri] - 128, Y[tri] - 128], -1).astype(np.int32)
    for p in pts:
        cv2.fillConvexPoly(img, p, 1, lineType=cv2.LINE_8, shift=8)
    cvcov = img > 0
    inter, union = (cov & cvcov).sum(), (cov | cvcov).sum()
    assert cov.sum() > 1200 and inter / union > 0.90, inter / union
    # every disagreement lies on the silhouette ring: eroding the OpenCV fill by one pixel puts it inside ours, dilating ours covers it
    k = np.ones((3, 3), np.uint8)
    assert not (cv2.erode(img, k).astype(bool) & ~cov).any()
    assert not (cvcov & ~cv2.dilate(cov.astype(np.uint8), k).astype(bool)).any()


def test_antialiased_mask_equals_area_coverage_on_axis_aligned_edges():
    """Closed form for the restated dr.antialias: for a straight, axis-aligned silhouette edge the blend weight is
    the distance of the edge from the pixel-pair midpoint, so the antialiased mask equals the exact fraction of the
    pixel that the shape covers."""
    H, W = 56, 80
    P = refpath.projection_matrix(90.0, 90.0, 40.0, 28.0, W, H).astype(np.float64)
    d = 3.0
    x0, x1, y0, y1 = 20.3, 60.7, 10.4, 40.6  # edges in pixel units

    def cam(px, py):
        nx, ny = px / W * 2 - 1, py / H * 2 - 1
        return [(nx + P[0, 2]) * d / P[0, 0], (ny + P[1, 2]) * d / P[1, 1], 0.0]

    v = np.array([cam(x0, y0), cam(x1, y0), cam(x1, y1), cam(x0, y1)], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]])
    mesh = refpath.Mesh(v, f, vtx_color=np.ones((4, 3), np.float32))
    q = np.array([[0.0, 0.0, 0.0, 1.0]], dtype=np.float32)
    t = np.array([[0.0, 0.0, -d]], dtype=np.float32)
    m = refpath.render(mesh, P, torch.from_numpy(q), torch.from_numpy(t), H, W)["mask"][0, ..., 0].numpy()
    rows, cols = slice(13, 38), slice(23, 58)  # away from the corners
    assert np.allclose(m[rows, 20], 0.7, atol=2e-3) and np.allclose(m[rows, 19], 0.0, atol=1e-6)    # left edge at 20.3
    assert np.allclose(m[rows, 60], 0.7, atol=2e-3) and np.allclose(m[rows, 61], 0.0, atol=1e-6)    # right edge at 60.7
    assert np.allclose(m[10, cols], 0.6, atol=2e-3) and np.allclose(m[9, cols], 0.0, atol=1e-6)     # bottom edge at 10.4
    assert np.allclose(m[40, cols], 0.6, atol=2e-3) and np.allclose(m[41, cols], 0.0, atol=1e-6)    # top edge at 40.6
    assert np.allclose(m[rows, cols], 1.0, atol=1e-6)
    # total mask mass = area of the rectangle up to the four corner pixels
    assert abs(m.sum() - (x1 - x0) * (y1 - y0)) < 1.0


# ----------------------------------------------------------------------------------------------
# golden vectors produced by the REFERENCE's own functions (tests/golden/make_reference_vectors.py imports
# /root/reference/diffdope/diffdope.py in the build container and calls the parts that need no nvdiffrast)

REFVEC = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz")


@pytest.fixture(scope="module")
def refvec():
    return np.load(REFVEC)


def test_reference_vectors_pose_and_projection(refvec):
    g = refvec
    # oracle: torch restatement and the canonical float32 version the kernels mirror
    m = refpath.matrix_batch_44_from_position_quat(torch.from_numpy(g["pose_q"]), torch.from_numpy(g["pose_p"])).numpy()
    assert np.array_equal(m, g["pose_mtx"])
    _, M = nvdr.canonical_pose(g["pose_q"], g["pose_p"])  # renormalises q (already unit): 1 ulp apart at most
    assert np.abs(M - g["pose_mtx"]).max() <= 3e-7
    for c, P in zip(g["cam_params"], g["cam_proj"]):
        assert np.array_equal(refpath.projection_matrix(c[0], c[1], c[2], c[3], int(c[4]), int(c[5])), P)
    # product-side mirrors of the same reference functions (pure torch, no GPU needed)
    import diffdope as dd

    m2 = dd.matrix_batch_44_from_position_quat(torch.from_numpy(g["pose_q"]), torch.from_numpy(g["pose_p"])).numpy()
    assert np.array_equal(m2, g["pose_mtx"])
    for c, P in zip(g["cam_params"], g["cam_proj"]):
        cam = dd.Camera(fx=c[0], fy=c[1], cx=c[2], cy=c[3], im_width=int(c[4]), im_height=int(c[5]))
        assert np.array_equal(cam.get_projection_matrix().numpy(), P)
    assert np.array_equal(np.array([refpath.lr_schedule(it, 60, 20, 0.1) for it in range(61)]), g["sched"])


def test_reference_vectors_losses_values_logs_and_gradients(refvec):
    g = refvec
    B = g["loss_lr"].shape[0]
    w = g["loss_weights"]
    cfg = dict(l1_rgb_with_mask=True, weight_rgb=float(w[0]), l1_depth_with_mask=True, weight_depth=float(w[1]), l1_mask=True, weight_mask=float(w[2]))
    rgb = torch.tensor(g["loss_rgb"], requires_grad=True)
    depth = torch.tensor(g["loss_depth"], requires_grad=True)
    mask = torch.tensor(g["loss_mask"], requires_grad=True)
    gt = {"rgb": torch.from_numpy(g["loss_gt_rgb"][0]), "depth": torch.from_numpy(g["loss_gt_depth"][0]), "segmentation": torch.from_numpy(g["loss_gt_seg"][0])}
    # the reference compares against per-hypothesis targets; the fixture's rgb / depth targets differ per hypothesis, so
    # evaluate the oracle one hypothesis at a time with the batch mean's 1/B restored
    total = torch.zeros(1)
    logged = {"rgb": [], "depth": [], "mask_selection": []}
    for b in range(B):
        gt_b = {"rgb": torch.from_numpy(g["loss_gt_rgb"][b]), "depth": torch.from_numpy(g["loss_gt_depth"][b]), "segmentation": torch.from_numpy(g["loss_gt_seg"][b])}
        t_b, l_b = refpath.losses({"rgb": rgb[b:b + 1], "depth": depth[b:b + 1], "mask": mask[b:b + 1]}, gt_b, torch.from_numpy(g["loss_lr"][b:b + 1]), cfg)
        total = total + t_b / B
        for k in logged:
            logged[k].append(float(l_b[k][0]))
    total.backward()
    assert np.allclose(float(total), g["loss_values"].sum(), rtol=1e-6)
    assert np.allclose(logged["rgb"], g["logged_rgb"], rtol=1e-6) and np.allclose(logged["depth"], g["logged_depth"], rtol=1e-6)
    assert np.allclose(logged["mask_selection"], g["logged_mask"], rtol=1e-6)
    assert np.allclose(rgb.grad.numpy(), g["grad_rgb"], rtol=1e-5, atol=1e-9)
    assert np.allclose(depth.grad.numpy(), g["grad_depth"], rtol=1e-5, atol=1e-9)
    assert np.allclose(mask.grad.numpy(), g["grad_mask"], rtol=1e-5, atol=1e-9)
    # the product's torch-written loss functions (the autograd path for user losses) against the same vectors
    import types

    import diffdope as dd

    class Mock:
        pass

    d = Mock()
    r2, d2, m2 = torch.tensor(g["loss_rgb"], requires_grad=True), torch.tensor(g["loss_depth"], requires_grad=True), torch.tensor(g["loss_mask"], requires_grad=True)
    d.renders = {"rgb": r2, "depth": d2, "mask": m2}
    d.gt_tensors = {"rgb": torch.from_numpy(g["loss_gt_rgb"]), "depth": torch.from_numpy(g["loss_gt_depth"]), "segmentation": torch.from_numpy(g["loss_gt_seg"])}
    d.learning_rates = torch.from_numpy(g["loss_lr"])
    d.cfg = types.SimpleNamespace(losses=types.SimpleNamespace(weight_rgb=float(w[0]), weight_depth=float(w[1]), weight_mask=float(w[2])))
    d.optimization_results = [{}]
    d.logged = {}
    d.add_loss_value = lambda key, values, values_weighted=None: d.logged.__setitem__(key, values.detach().numpy())
    vals = [dd.l1_rgb_with_mask(d), dd.l1_depth_with_mask(d), dd.l1_mask(d)]
    sum(vals).backward()
    assert np.allclose([float(v) for v in vals], g["loss_values"], rtol=1e-6)
    assert np.allclose(d.logged["rgb"], g["logged_rgb"], rtol=1e-6) and np.allclose(d.logged["mask_selection"], g["logged_mask"], rtol=1e-6)
    assert np.allclose(r2.grad.numpy(), g["grad_rgb"], rtol=1e-5, atol=1e-9) and np.allclose(m2.grad.numpy(), g["grad_mask"], rtol=1e-5, atol=1e-9)
    # find_crop
    m = torch.zeros(40, 60, 3)
    y0, y1, x0, x1 = g["crop_mask_box"]
    m[y0:y1, x0:x1] = 1.0
    assert list(dd.find_crop(m)) == list(g["crop"])


def test_reference_vectors_image_loading(refvec):
    """`Image.__post_init__` of the reference (cv2 pipeline: BGR->RGB, /255, flip, resize, depth / 100) against the
    loaders used by the tests (scene_util.load_image) and by the product (`diffdope.Image`)."""
    import diffdope as dd

    g = refvec
    data = os.path.join(su.ROOT, "data", "example", "scene")
    for key, fname, depth in (("rgb", "rgb.png", False), ("depth", "depth.png", True), ("seg", "seg.png", False)):
        ours = su.load_image(os.path.join(data, fname), 0.5, depth=depth)
        prod = dd.Image(img_path=os.path.join(data, fname), img_resize=0.5, depth=depth).img_tensor.numpy()
        for im in (ours, prod):
            assert list(im.shape) == list(g["img_%s_shape" % key])
            assert np.allclose(float(im.astype(np.float64).sum()), float(g["img_%s_sum" % key]), rtol=1e-7)
            assert np.array_equal(im[::37, ::41], g["img_%s_sample" % key])


REFRUN = os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz")


def test_oracle_reproduces_the_reference_loop_run_on_cpu():
    """tests/golden/reference_run.npz is the output of the reference's own, unmodified `DiffDope.run_optimization`
    (4 iterations, 2 hypotheses, rgb + depth + mask losses, quarter resolution) executed on the CPU with the four
    nvdiffrast ops served by the oracle's restatements (make_reference_run.py). The oracle's restatement of
    everything around those ops -- render-graph wiring, depth sign, background masking, loss structure, logging keys,
    schedule over nb_iterations+1 steps, SGD on the raw parameters, argmin, get_pose -- must reproduce it."""
    g = np.load(REFRUN)
    assert list(g["loss_keys"]) == ["rgb", "depth", "mask_selection"]
    arr = su.example_mesh_arrays()
    gt = {k: torch.from_numpy(v) for k, v in su.example_targets(float(g["resize"])).items()}
    H, W = gt["rgb"].shape[:2]
    q, t = su.example_pose()
    pose0 = g["pose0"]
    assert np.allclose(pose0[:, :4], q, atol=1e-6) and np.allclose(pose0[:, 4:], t, atol=1e-6)
    assert np.allclose(g["lr"], su.lr_multipliers(2, 0.05, 0.5))
    cfg = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
    hyper = dict(nb_iterations=3, base_lr=20.0, lr_decay=0.1, learning_rate_base=1)
    for cull in (False, True):  # a GL context does not cull; the back-face rule must not change anything here
        mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
        mesh.cull = cull
        o = refpath.run_optimization(mesh, su.projection(), pose0[:, :4], pose0[:, 4:], gt, g["lr"], cfg, hyper, H, W)
        for k in ("rgb", "depth", "mask_selection"):
            assert o["losses"][k].shape == g["loss_" + k].shape == (4, 2)
            assert np.allclose(o["losses"][k], g["loss_" + k], rtol=2e-6, atol=1e-10), k
        assert np.abs(o["final"] - g["final"]).max() < 2e-6
        assert np.abs(o["mtx"] - g["mtx"]).max() < 2e-6
        assert refpath.argmin_hypothesis(o["losses"]) == int(g["argmin"])
        assert np.abs(o["mtx"][-1][int(g["argmin"])] - g["best_pose"]).max() < 2e-6
    # render conventions of the first iteration: background rgb 0, background depth = -t_z, mask in [0,1]
    r = refpath.render(refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"]), su.projection(), torch.from_numpy(pose0[:, :4]),
                       torch.from_numpy(pose0[:, 4:]), H, W)
    assert np.allclose(r["rgb"].numpy()[:, ::7, ::9], g["rgb0_sample"], atol=1e-6)
    assert np.allclose(r["depth"].numpy()[:, ::7, ::9], g["depth0_sample"], atol=1e-5)
    assert np.isclose(float(r["rgb"].double().sum()), float(g["rgb_sum"][0]), rtol=1e-6)
    assert np.isclose(float(r["depth"].double().sum()), float(g["depth_sum"][0]), rtol=1e-6)


def _cube_reference_scenario():
    """Inputs of the untextured scenario of tests/golden/make_reference_run.py, as the reference's loaders produce them."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("cube_scenario", os.path.join(os.path.dirname(__file__), "golden", "cube_scenario.py"))
    mrr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mrr)
    rgb, depth, seg = mrr.cube_targets()
    gt = dict(rgb=np.ascontiguousarray((rgb / 255.0)[::-1], dtype=np.float32), depth=np.ascontiguousarray((depth / 100)[::-1], dtype=np.float32),
              segmentation=np.ascontiguousarray((seg / 255.0)[::-1], dtype=np.float32))
    pos = (mrr.CUBE_V.astype(np.float32) * np.float32(0.01)).astype(np.float32)
    col = (mrr.CUBE_C[..., :3] / 255.0).astype(np.float32)
    c = mrr.CUBE_CAM
    P = refpath.projection_matrix(c["fx"], c["fy"], c["cx"], c["cy"], c["im_width"], c["im_height"])
    return dict(pos=pos, tri=mrr.CUBE_F, col=col, gt=gt, P=P, q=mrr.CUBE_Q.astype(np.float32), t=mrr.CUBE_T.astype(np.float32), H=mrr.CUBE_HW[0], W=mrr.CUBE_HW[1])


def test_oracle_reproduces_the_reference_loop_untextured_branch():
    """Same as above for the vertex-colour branch of the reference (`diffdope.py:229-231,1677-1686`, Mesh loader's else
    branch), with the default loss configuration (mask only) and with all three losses; 3 hypotheses x 3 iterations."""
    g = np.load(REFRUN)
    s = _cube_reference_scenario()
    gt = {k: torch.from_numpy(v) for k, v in s["gt"].items()}
    mesh = refpath.Mesh(s["pos"], s["tri"], vtx_color=s["col"])
    mesh.cull = False
    for tag, cfg in (("cube_mask", dict(l1_mask=True, weight_mask=1.0)),
                     ("cube_all", dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0))):
        assert list(g[tag + "_keys"]) == list(k for k in ("rgb", "depth", "mask_selection") if ("l1_rgb_with_mask" in cfg or k == "mask_selection"))
        lr = g[tag + "_lr"]
        assert np.allclose(lr, su.lr_multipliers(3, 0.05, 0.5, seed=1))
        hyper = dict(nb_iterations=2, base_lr=20.0, lr_decay=0.1, learning_rate_base=1)
        o = refpath.run_optimization(mesh, s["P"], np.tile(s["q"], (3, 1)), np.tile(s["t"], (3, 1)), gt, lr, cfg, hyper, s["H"], s["W"])
        for k in g[tag + "_keys"]:
            assert np.allclose(o["losses"][k], g[tag + "_loss_" + k], rtol=3e-6, atol=1e-10), (tag, k)
        assert np.abs(o["final"] - g[tag + "_final"]).max() < 3e-6
        assert refpath.argmin_hypothesis(o["losses"]) == int(g[tag + "_argmin"])
    r = refpath.render(mesh, s["P"], torch.from_numpy(np.tile(s["q"], (3, 1))), torch.from_numpy(np.tile(s["t"], (3, 1))), s["H"], s["W"])
    assert np.allclose(r["rgb"].numpy()[:, ::5, ::7], g["cube_all_rgb0_sample"], atol=1e-6)


def test_reference_vectors_visualisation_helpers(refvec):
    """`make_grid` / `make_grid_overlay_batch` (`diffdope.py:337-528`, the image side of render_img / make_animation):
    the product's helpers produce the reference's images bit for bit on the same batch (grid layout, alpha overlay,
    contour, flip, resize)."""
    import diffdope as dd

    g = refvec
    fg, bg = torch.from_numpy(g["viz_fg"]), torch.from_numpy(g["viz_bg"])
    a = dd.make_grid_overlay_batch(background=bg, foreground=fg, alpha=0.7, row=2, final_width=300, add_background=True, add_contour=True,
                                   color_countour=[0.46, 0.73, 0], flip_result=True)
    assert a.dtype == np.uint8 and np.array_equal(a, g["viz_overlay"])
    b = dd.make_grid_overlay_batch(background=bg, foreground=fg, alpha=0.5, row=3, final_width=200, add_background=False, add_contour=True, flip_result=False)
    assert np.array_equal(b, g["viz_overlay_plain"])
    assert np.array_equal(dd.make_grid(fg.permute(0, 3, 1, 2), nrow=2).numpy(), g["viz_grid"])
    # the reference crashes with add_contour=False (alpha_img unbound, diffdope.py:493-514); the product returns the plain overlay
    c = dd.make_grid_overlay_batch(background=bg, foreground=fg, alpha=0.5, row=3, final_width=200, add_background=True, add_contour=False, flip_result=False)
    assert c.shape == b.shape


def test_textured_interior_gradient_matches_finite_differences():
    """The texture -> interpolate -> rasterize -> xfm backward chain of the oracle against central finite differences, on
    a smooth texture and a loss restricted to interior pixels (3 px away from the silhouette at every probed pose), where
    the render is a smooth function of the pose. All seven pose parameters."""
    v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0], [0, 0, 0.3]], dtype=np.float32)
    f = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]])  # a shallow pyramid: four faces, non-trivial barycentrics
    uv = np.array([[0.1, 0.1], [0.9, 0.1], [0.9, 0.9], [0.1, 0.9], [0.5, 0.5]], dtype=np.float32)
    ty, tx = np.mgrid[0:64, 0:64] / 64.0
    tex = np.stack([0.5 + 0.4 * np.sin(5 * tx + 2 * ty), 0.5 + 0.4 * np.cos(4 * ty - tx), 0.5 + 0.3 * np.sin(3 * tx * ty + 1)], -1).astype(np.float32)
    mesh = refpath.Mesh(v, f, uv=uv, tex=tex)
    H, W = 72, 96
    P = refpath.projection_matrix(130.0, 130.0, 48.0, 36.0, W, H)
    q0 = np.array([[0.12, -0.09, 0.05, 0.985]], dtype=np.float64)
    t0 = np.array([[0.05, -0.03, -5.0]], dtype=np.float64)
    base = refpath.render(mesh, P, torch.tensor(q0, dtype=torch.float32), torch.tensor(t0, dtype=torch.float32), H, W)
    cov = base["rast_out"][0, ..., 3].numpy() > 0
    import cv2

    interior = cv2.erode(cov.astype(np.uint8), np.ones((9, 9), np.uint8)).astype(bool)  # 4 px inside the silhouette
    assert interior.sum() > 600
    yy, xx = np.mgrid[0:H, 0:W]
    wgt = torch.tensor((interior * (1.0 + 0.5 * np.sin(xx / 6.0) * np.cos(yy / 5.0)))[..., None].astype(np.float32) * np.array([1.0, -0.7, 0.4], np.float32))

    def L(q, t, grad=False):
        qq = torch.tensor(q, dtype=torch.float32, requires_grad=True)
        tt = torch.tensor(t, dtype=torch.float32, requires_grad=True)
        r = refpath.render(mesh, P, qq, tt, H, W)
        l = (r["rgb"][0] * wgt).sum() + 0.3 * (r["depth"][0] * wgt[..., 0]).sum()
        if grad:
            l.backward()
            return float(l.detach()), np.concatenate([qq.grad.numpy()[0], tt.grad.numpy()[0]])
        return float(l.detach())

    _, g = L(q0, t0, True)
    h = 2e-3
    fd = []
    for i in range(7):
        qp, qm, tp, tm = q0.copy(), q0.copy(), t0.copy(), t0.copy()
        (qp if i < 4 else tp)[0, i % 4 if i < 4 else i - 4] += h
        (qm if i < 4 else tm)[0, i % 4 if i < 4 else i - 4] -= h
        fd.append((L(qp, tp) - L(qm, tm)) / (2 * h))
    fd = np.array(fd)
    # the texture is piecewise bilinear: allow 3 % of the largest component
    assert np.abs(g - fd).max() < 0.03 * np.abs(fd).max(), (g, fd)


def test_antialias_gradient_closed_form_centroid_shift():
    """The silhouette gradient (restated AntialiasGradKernel + everything upstream of it) on a fronto-parallel rectangle:
    sum_px x*mask = area * centroid_x in the continuous picture, so d/dt_x ~ area * fx / depth pixels per unit and
    sum_px mask (= area) is invariant under a sideways translation. The analytic gradient must agree with a central
    finite difference of the oracle's own forward pass and with the continuous value, both to 3 % of that value (the
    remainder comes from the four corner pixels, where the pairwise blend is neither an exact area nor exactly
    differentiated by the published gradient kernel)."""
    H, W = 56, 80
    fx = fy = 90.0
    P = refpath.projection_matrix(fx, fy, 40.0, 28.0, W, H).astype(np.float64)
    d = 3.0
    x0, x1, y0, y1 = 20.3, 60.7, 10.4, 40.6

    def cam(px, py):
        nx, ny = px / W * 2 - 1, py / H * 2 - 1
        return [(nx + P[0, 2]) * d / P[0, 0], (ny + P[1, 2]) * d / P[1, 1], 0.0]

    v = np.array([cam(x0, y0), cam(x1, y0), cam(x1, y1), cam(x0, y1)], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]])
    mesh = refpath.Mesh(v, f, vtx_color=np.ones((4, 3), np.float32))
    yy, xx = np.mgrid[0:H, 0:W]
    area = (x1 - x0) * (y1 - y0)
    scale = area * fx / d

    def value(wimg, tx, ty):
        q = torch.tensor([[0.0, 0.0, 0.0, 1.0]])
        t = torch.tensor([[tx, ty, -d]], dtype=torch.float32)
        return float((refpath.render(mesh, P, q, t, H, W)["mask"][0, ..., 0] * wimg).sum())

    for wnp, ideal in ((xx + 0.5, [scale, 0.0]), (yy + 0.5, [0.0, scale]), (np.ones((H, W)), [0.0, 0.0])):
        wimg = torch.tensor(wnp.astype(np.float32))
        q = torch.tensor([[0.0, 0.0, 0.0, 1.0]], requires_grad=True)
        t = torch.tensor([[0.0, 0.0, -d]], requires_grad=True)
        (refpath.render(mesh, P, q, t, H, W)["mask"][0, ..., 0] * wimg).sum().backward()
        got = t.grad.numpy()[0, :2]
        h = 2e-3
        fd = np.array([(value(wimg, h, 0.0) - value(wimg, -h, 0.0)) / (2 * h), (value(wimg, 0.0, h) - value(wimg, 0.0, -h)) / (2 * h)])
        assert np.allclose(got, fd, rtol=1e-3, atol=0.03 * scale), (got, fd)
        assert np.allclose(got, ideal, rtol=0.03, atol=0.03 * scale), (got, ideal)
