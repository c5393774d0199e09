"""Size-independent properties of the CPU oracle (oracle/nvdr.py, oracle/refpath.py) on seeded random inputs:
the raster rule is watertight and order-independent, the texture lookup is periodic and convex, the
antialiased mask only differs from coverage next to a silhouette, and a pose that rendered its own target
has zero loss and zero gradient. The same properties are asserted for the CUDA path at full size in
tests/test_gpu_configs.py; here they pin the checker itself (no GPU)."""
import numpy as np
import torch

from oracle import nvdr, refpath

F = np.float32


def _random_convex_fan(rng, n, W, H):
    """A convex polygon with n corners at random sub-pixel positions, fanned from a random interior point:
    the n triangles partition it."""
    ang = np.sort(rng.uniform(0, 2 * np.pi, n))
    ang += np.linspace(0, 1e-3, n)  # distinct
    r = rng.uniform(0.55, 0.8)
    ring = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    centre = ring.mean(0) + rng.uniform(-0.05, 0.05, 2)
    xy = np.concatenate([centre[None], ring], 0)
    w = rng.uniform(0.8, 1.6, (n + 1, 1))
    z = rng.uniform(-0.3, 0.3, (n + 1, 1))
    clip = np.concatenate([xy * w, z * w, w], 1).astype(F)[None]
    tri = np.array([[0, 1 + k, 1 + (k + 1) % n] for k in range(n)])
    return clip, tri


def test_raster_partition_of_a_convex_polygon_is_watertight():
    """Each pixel centre inside a fanned convex polygon is covered by exactly one of its triangles, whatever the
    sub-pixel positions and windings: the union equals the polygon rasterised as separate single triangles summed."""
    rng = np.random.default_rng(7)
    H, W = 48, 64
    for trial in range(6):
        clip, tri = _random_convex_fan(rng, int(rng.integers(5, 12)), W, H)
        if trial % 2:
            tri = tri[:, ::-1].copy()
        count = np.zeros((H, W), int)
        for k in range(tri.shape[0]):
            count += nvdr.rasterize(clip, tri[k:k + 1], H, W)[0, ..., 3] > 0
        assert count.max() == 1, "a pixel centre was claimed by two triangles of a partition"
        union = nvdr.rasterize(clip, tri, H, W)[0, ..., 3] > 0
        assert np.array_equal(union, count == 1)
        # no holes: every row of the covered set is one interval (the polygon is convex)
        for y in range(H):
            xs = np.nonzero(union[y])[0]
            if xs.size:
                assert xs[-1] - xs[0] + 1 == xs.size, "hole inside a convex polygon"
        assert union.sum() > 300


def test_raster_result_does_not_depend_on_triangle_order():
    """Random overlapping triangles: permuting the index buffer changes the ids but not which triangle (as a vertex
    triple) wins each pixel, nor its z/w and barycentrics (depth ties have probability zero here)."""
    rng = np.random.default_rng(11)
    H, W = 40, 40
    V, T = 30, 24
    w = rng.uniform(0.7, 2.0, (V, 1))
    clip = np.concatenate([rng.uniform(-1, 1, (V, 2)) * w, rng.uniform(-0.9, 0.9, (V, 1)) * w, w], 1).astype(F)[None]
    tri = np.stack([rng.choice(V, 3, replace=False) for _ in range(T)])
    perm = rng.permutation(T)
    r0 = nvdr.rasterize(clip, tri, H, W)[0]
    r1 = nvdr.rasterize(clip, tri[perm], H, W)[0]
    cov = r0[..., 3] > 0
    assert np.array_equal(cov, r1[..., 3] > 0) and cov.sum() > 200
    id0 = r0[..., 3][cov].astype(int) - 1
    id1 = r1[..., 3][cov].astype(int) - 1
    assert np.array_equal(tri[id0], tri[perm][id1])
    assert np.array_equal(r0[..., :3][cov], r1[..., :3][cov])


def test_texture_lookup_is_periodic_and_a_convex_combination():
    rng = np.random.default_rng(3)
    tex = rng.random((8, 16, 3)).astype(F)
    # uv on a 1/64 grid: u + k is exact in float32, so wrap-mode periodicity is bitwise
    uv = (rng.integers(-64, 128, (1, 5, 7, 2)) / 64.0).astype(F)
    base = nvdr.texture_linear(tex, uv)
    for du, dv in ((1, 0), (0, -2), (3, 5)):
        shifted = nvdr.texture_linear(tex, (uv + np.array([du, dv], dtype=F)).astype(F))
        assert np.array_equal(base, shifted)
    assert base.min() >= tex.min() - 1e-6 and base.max() <= tex.max() + 1e-6
    const = np.full((4, 4, 3), 0.375, dtype=F)
    assert np.array_equal(nvdr.texture_linear(const, rng.uniform(-3, 3, (1, 4, 4, 2)).astype(F)), np.full((1, 4, 4, 3), 0.375, dtype=F))


def _cube():
    v = np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)], dtype=F)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tri = np.array([t for a, b, c, d in quads for t in ((a, b, c), (a, c, d))])
    col = (v + 0.5).astype(F)
    return v, tri, col


def _render_cube(q, t, H=40, W=56):
    v, tri, col = _cube()
    mesh = refpath.Mesh(v, tri, vtx_color=col)
    proj = refpath.projection_matrix(60.0, 60.0, W / 2 - 0.3, H / 2 + 0.2, W, H)
    r = refpath.render(mesh, proj, torch.tensor(q[None]), torch.tensor(t[None]), H, W)
    return mesh, proj, {k: v_.detach() for k, v_ in r.items()}


def test_antialiased_mask_differs_from_coverage_only_next_to_a_silhouette():
    q = np.array([0.21, -0.35, 0.12, 0.9], dtype=F)
    t = np.array([0.1, -0.05, -3.0], dtype=F)
    _, _, r = _render_cube(q, t)
    cov = (r["rast_out"][0, ..., 3] > 0).numpy()
    mask = r["mask"][0].numpy()
    assert cov.sum() > 150
    assert np.array_equal(mask[..., 0], mask[..., 1]) and np.array_equal(mask[..., 0], mask[..., 2])
    m = mask[..., 0]
    assert m.min() >= 0.0 and m.max() <= 1.0 + 1e-6
    # pixels whose 4-neighbourhood has the same coverage keep the plain coverage value (the oracle interpolates the
    # constant 1 like the reference does, diffdope.py:212-213: u + v + (1-u-v) = 1 within an ulp, and blends between
    # two covered pixels move it by alpha * (1 - 1))
    pad = np.pad(cov, 1, mode="edge")
    same = (pad[1:-1, :-2] == cov) & (pad[1:-1, 2:] == cov) & (pad[:-2, 1:-1] == cov) & (pad[2:, 1:-1] == cov)
    assert np.allclose(m[same], cov[same].astype(F), rtol=0, atol=1e-6)
    changed = np.abs(m - cov.astype(F)) > 1e-6
    assert changed.sum() > 10, "the silhouette ring must carry fractional coverage"
    # fractional values lower covered pixels and raise uncovered ones, never the other way round
    assert np.all(m[changed & cov] < 1.0) and np.all(m[changed & ~cov] > 0.0)


def test_zero_loss_and_zero_gradient_at_the_pose_that_rendered_the_target():
    q = np.array([0.21, -0.35, 0.12, 0.9], dtype=F)
    t = np.array([0.1, -0.05, -3.0], dtype=F)
    H, W = 40, 56
    mesh, proj, r = _render_cube(q, t, H, W)
    gt = {"rgb": r["rgb"][0], "depth": r["depth"][0], "segmentation": r["mask"][0]}
    cfg = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
    logged, gq, gt_, _ = refpath.forward_backward(mesh, proj, q[None], t[None], gt, np.array([3.0], dtype=F), cfg, H, W)
    for k in ("rgb", "depth", "mask_selection"):
        assert float(logged[k][0]) == 0.0
    assert np.all(gq == 0.0) and np.all(gt_ == 0.0)  # abs'(0) = 0
    # and a displaced pose has positive losses and a non-zero gradient
    logged, gq, gt_, _ = refpath.forward_backward(mesh, proj, q[None], (t + np.array([0.05, 0, 0], dtype=F))[None], gt, np.array([3.0], dtype=F), cfg, H, W)
    assert all(float(logged[k][0]) > 0 for k in ("rgb", "depth", "mask_selection"))
    assert np.abs(gt_).max() > 0 and np.abs(gq).max() > 0


def test_hypotheses_are_independent_and_shards_reassemble():
    """Loss values and gradients of a hypothesis do not depend on which other hypotheses share its batch once the
    global batch size is the divisor: the property the multi-GPU sharding rests on."""
    rng = np.random.default_rng(5)
    q = np.array([0.21, -0.35, 0.12, 0.9], dtype=F)
    t = np.array([0.1, -0.05, -3.0], dtype=F)
    H, W = 32, 40
    mesh, proj, r = _render_cube(q, t, H, W)
    gt = {"rgb": r["rgb"][0], "depth": r["depth"][0], "segmentation": r["mask"][0]}
    B = 4
    qs = (q[None] + rng.normal(0, 0.03, (B, 4))).astype(F)
    ts = (t[None] + rng.normal(0, 0.03, (B, 3))).astype(F)
    lr = rng.uniform(0.1, 5.0, B).astype(F)
    cfg = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
    logged, gq, gtr, _ = refpath.forward_backward(mesh, proj, qs, ts, gt, lr, cfg, H, W)
    for lo, hi in ((0, 1), (1, 4)):
        l2, gq2, gt2, _ = refpath.forward_backward(mesh, proj, qs[lo:hi], ts[lo:hi], gt, lr[lo:hi], cfg, H, W, b_global=B)
        for k in logged:
            assert np.array_equal(logged[k][lo:hi].numpy(), l2[k].numpy())
        assert np.allclose(gq[lo:hi], gq2, rtol=1e-6, atol=1e-12) and np.allclose(gtr[lo:hi], gt2, rtol=1e-6, atol=1e-12)
