"""GPU parity of the north_star extensions that have no reference counterpart:
Sobel-edge loss, linear-mipmap-linear texture filter, Adam pose step. Each is checked through the C ABI
against its own oracle (oracle/refpath.py, oracle/nvdr.py); defaults stay the reference's behaviour."""
import os

import numpy as np
import pytest
import torch

import scene_util as su
from test_gpu_parity import ALL, Example, _angle_deg, _cfg, _loss_table

pytestmark = pytest.mark.gpu

EDGE_ONLY = dict(l1_edge=True, weight_edge=1.0)
FULL = dict(ALL, l1_edge=True, weight_edge=0.5)


@pytest.fixture(scope="module")
def ex_q():
    return Example(0.5)


def _oracle(ex, qs, ts, lr, losses, mip=False, window=None, b_global=None):
    from oracle import refpath

    mesh = ex.oracle_mesh()
    if mip:
        mesh.texture_filter = "linear-mipmap-linear"
    return refpath.forward_backward(mesh, ex.P, qs, ts, ex.gt_t(), lr, losses, ex.H, ex.W, window=window, b_global=b_global)


@pytest.mark.parametrize("losses", [EDGE_ONLY, FULL])
def test_edge_loss_and_gradient_match_oracle(ex_q, losses):
    ex = ex_q
    B = 3
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = su.lr_multipliers(B)
    ex.sc.set_texture_filter("linear")
    loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, losses))
    logged, gq, gtr, _ = _oracle(ex, qs, ts, lr, losses)
    assert logged["edge"].numpy().min() > 1e-5, "the edge loss must be live"
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    assert np.abs(go).max() > 0
    assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()


def test_edge_loss_window_zero_padding_and_shard(ex_q):
    """The Sobel stencil is zero-padded at the loss window (not at the frame), for render and target alike;
    a 2-of-5 shard carries the global divisor."""
    ex = ex_q
    B = 2
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = su.lr_multipliers(B)
    seg = ex.gt["segmentation"]
    ys, xs = np.nonzero(seg[..., 0] > 0)
    # a window that cuts through the object: its border lies on covered pixels
    cy, cx = int((ys.min() + ys.max()) // 2), int((xs.min() + xs.max()) // 2)
    win = (cy - 20, cx - 90, 70, 100)
    try:
        ex.sc.set_window(*win)
        loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, FULL), b_global=5)
        logged, gq, gtr, _ = _oracle(ex, qs, ts, lr, FULL, window=win, b_global=5)
        assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
        go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
        assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
    finally:
        ex.sc.set_window(0, 0, ex.H, ex.W)


def test_mipmap_render_loss_gradient_match_oracle(ex_q):
    from oracle import refpath

    ex = ex_q
    B = 3
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = su.lr_multipliers(B)
    try:
        ex.sc.set_texture_filter("linear-mipmap-linear")
        out = ex.sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda())
        mesh = ex.oracle_mesh()
        mesh.texture_filter = "linear-mipmap-linear"
        r = refpath.render(mesh, ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W)
        rgb_o, rgb_g = r["rgb"].numpy(), out["rgb"].cpu().numpy()
        assert np.array_equal(r["rast_out"].numpy()[..., 3], out["rast"].cpu().numpy()[..., 3])
        assert np.abs(rgb_o - rgb_g).max() <= 1e-4, "trilinear colour within 1e-4 (float32 level of detail vs float64 in the oracle)"
        # the filter must actually do something on this heavily minified texture
        lin = refpath.render(ex.oracle_mesh(), ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W)["rgb"].numpy()
        assert np.abs(lin - rgb_o).max() > 0.05
        for losses in (ALL, FULL):
            loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, losses))
            logged, gq, gtr, _ = _oracle(ex, qs, ts, lr, losses, mip=True)
            assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
            go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
            su.record("mipmap_gradient_%d_losses" % len([k for k in losses if k.startswith("l1_")]), grad_rel_err=su.grad_rel_err(go, gg))
            assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
    finally:
        ex.sc.set_texture_filter("linear")
    # back on the reference filter the render is bit-exact again
    out = ex.sc.render(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), want=("rgb",))
    assert np.array_equal(lin, out["rgb"].cpu().numpy())


def test_adam_trajectory_matches_torch_adam(ex_q):
    """ddope_optimize with DDOPE_OPT_ADAM against torch.optim.Adam on the oracle graph, 8 iterations."""
    from oracle import refpath

    ex = ex_q
    B, iters = 2, 8
    qs, ts = np.tile(ex.q, (B, 1)), np.tile(ex.t, (B, 1))
    lr = np.array([0.5, 2.0], dtype=np.float32)
    hyper = dict(nb_iterations=iters - 1, base_lr=0.002, lr_decay=0.5, learning_rate_base=1, optimizer="adam", adam_beta1=0.9, adam_beta2=0.99, adam_eps=1e-8)
    o = refpath.run_optimization(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, ALL, hyper, ex.H, ex.W)
    sched = [refpath.lr_schedule(it, iters - 1, hyper["base_lr"], hyper["lr_decay"]) for it in range(iters)]
    try:
        ex.sc.set_optimizer("adam", beta1=0.9, beta2=0.99, eps=1e-8)
        qd, td = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
        ph, lh = ex.sc.optimize(qd, td, torch.from_numpy(lr).cuda(), sched, _cfg(ex.n, ALL))
        # the first Adam step moves every parameter by exactly lr_0 (bias-corrected m/sqrt(v) = sign g)
        step0 = np.abs(ph[1].cpu().numpy() - ph[0].cpu().numpy())
        assert np.allclose(step0, sched[0], rtol=2e-3)
        fin = np.concatenate([qd.cpu().numpy(), td.cpu().numpy()], 1)
        assert _angle_deg(fin[:, :4], o["final"][:, :4]).max() < 0.1
        assert np.abs(fin[:, 4:] - o["final"][:, 4:]).max() < 1e-3
        assert np.abs(ph.cpu().numpy()[:3] - o["poses"][:3]).max() < 2e-5
        # continuing from stored moments (step0 = 4) reproduces the one-call trajectory bit for bit
        qd2, td2 = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
        ex.sc.set_optimizer("adam", beta1=0.9, beta2=0.99, eps=1e-8, step0=0)
        ex.sc.optimize(qd2, td2, torch.from_numpy(lr).cuda(), sched[:4], _cfg(ex.n, ALL))
        ex.sc.set_optimizer("adam", beta1=0.9, beta2=0.99, eps=1e-8, step0=4)
        ex.sc.optimize(qd2, td2, torch.from_numpy(lr).cuda(), sched[4:], _cfg(ex.n, ALL))
        assert torch.equal(qd2, qd) and torch.equal(td2, td)
    finally:
        ex.sc.set_optimizer("sgd")


def test_extension_errors(ex_q):
    ex = ex_q
    with pytest.raises(RuntimeError):
        ex.sc.set_texture_filter("cubic")
    with pytest.raises(RuntimeError):
        ex.sc.set_optimizer("lbfgs")
    with pytest.raises(RuntimeError):
        ex.sc.set_optimizer("adam", beta1=1.5)
    # edge loss needs the rgb target
    n = ex.n
    sc = n.NativeScene(ex.arr["pos"], ex.arr["tri"], ex.arr["uv"], ex.arr["tex"])
    sc.set_camera(ex.P, ex.H, ex.W)
    sc.set_target(None, ex.g["depth"], ex.g["segmentation"])
    q = torch.from_numpy(ex.q[None]).cuda()
    t = torch.from_numpy(ex.t[None]).cuda()
    with pytest.raises(RuntimeError, match="edge loss needs the rgb target"):
        sc.loss_grad(q, t, torch.ones(1).cuda(), _cfg(n, EDGE_ONLY))


def test_parts_and_shards_are_bit_identical_with_extensions(ex_q):
    """16 hypotheses run as two internal parts (include/ddope_b200.h, "Streams"); four shards of 4 run unsplit.
    With Adam, the edge loss and the mipmap filter switched on, every table must still agree bit for bit: per-part
    offsets into the Adam moments, the history tables and the partial sums."""
    ex = ex_q
    B, iters = 16, 5
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = torch.from_numpy(su.lr_multipliers(B, 0.05, 2.0)).cuda()
    sched = [0.004 * 0.5 ** (i / (iters - 1) + 1) for i in range(iters)]
    cfg = _cfg(ex.n, FULL)
    try:
        ex.sc.set_optimizer("adam", beta1=0.9, beta2=0.99, eps=1e-8)
        ex.sc.set_texture_filter("linear-mipmap-linear")
        q, t = torch.from_numpy(qs).cuda().contiguous(), torch.from_numpy(ts).cuda().contiguous()
        ph, lh = ex.sc.optimize(q, t, lr, sched, cfg)
        assert ex.sc.last_launch_count() == 2 * (3 * iters + 1)
        for lo in range(0, B, 4):
            qk, tk = torch.from_numpy(qs[lo:lo + 4]).cuda().contiguous(), torch.from_numpy(ts[lo:lo + 4]).cuda().contiguous()
            pk, lk = ex.sc.optimize(qk, tk, lr[lo:lo + 4].contiguous(), sched, cfg, b_global=B)
            assert ex.sc.last_launch_count() == 3 * iters + 1
            assert torch.equal(pk, ph[:, lo:lo + 4]) and torch.equal(lk, lh[:, lo:lo + 4])
            assert torch.equal(qk, q[lo:lo + 4]) and torch.equal(tk, t[lo:lo + 4])
        assert float(lh[..., 3].min()) > 0 and not torch.equal(ph[0], ph[-1])
        # loss_grad splits too: its tables equal the first iteration of the optimisation
        loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), lr, cfg)
        assert torch.equal(loss, lh[0])
    finally:
        ex.sc.set_optimizer("sgd")
        ex.sc.set_texture_filter("linear")


def test_colour_attribute_gradients_match_oracle(ex_q, tmp_path):
    """`Mesh.enable_gradients_texture` (`diffdope/diffdope.py:909-920`, dead code in the reference): gradient of a loss on the
    rendered colour w.r.t. the texture texels (`ddope_render_bwd_attr`) against the oracle's dr.texture attribute backward, and
    w.r.t. the vertex colours of an untextured mesh against dr.interpolate's; then through the public API (`render_texture_batch`
    + autograd + an optimizer step that changes the texture)."""
    from oracle import refpath
    from test_gpu_parity import _cube_scene

    ex = ex_q
    B = 2
    qs, ts = su.perturbed_poses(ex.q, ex.t, B, seed=4, rot_deg=2.0, trans=0.02)
    rng = np.random.default_rng(0)
    w = rng.standard_normal((B, ex.H, ex.W, 3)).astype(np.float32)
    qn = qs / np.linalg.norm(qs, axis=1, keepdims=True)
    from diffdope import matrix_batch_44_from_position_quat

    mtx = matrix_batch_44_from_position_quat(torch.from_numpy(qn), torch.from_numpy(ts)).cuda().contiguous()
    g = ex.sc.render_attr_grad(mtx, torch.from_numpy(w).cuda())
    tex = torch.tensor(ex.arr["tex"], requires_grad=True)
    r = refpath.render(ex.oracle_mesh(), ex.P, torch.from_numpy(qs), torch.from_numpy(ts), ex.H, ex.W, tex=tex)
    (r["rgb"] * torch.from_numpy(w)).sum().backward()
    go, gg = tex.grad.numpy(), g.cpu().numpy()
    assert np.count_nonzero(go) > 1000 and np.array_equal(go != 0, gg != 0), "the same texels receive gradient"
    assert np.abs(go - gg).max() <= 1e-5 * np.abs(go).max()

    # vertex colours (cube)
    n, sc, mesh, P, gt, H, W = _cube_scene()
    qc = np.array([[0.3, 0.2, 0.1, 0.9], [0.0, 0.7, 0.1, 0.6]], dtype=np.float32)
    tc = np.array([[0.1, -0.05, -4.0], [-0.3, 0.2, -3.0]], dtype=np.float32)
    wc = rng.standard_normal((2, H, W, 3)).astype(np.float32)
    qcn = qc / np.linalg.norm(qc, axis=1, keepdims=True)
    mc = matrix_batch_44_from_position_quat(torch.from_numpy(qcn), torch.from_numpy(tc)).cuda().contiguous()
    gv = sc.render_attr_grad(mc, torch.from_numpy(wc).cuda()).cpu().numpy()
    vcol = torch.tensor(mesh.vtx_color, requires_grad=True)
    rc = refpath.render(mesh, P, torch.from_numpy(qc), torch.from_numpy(tc), H, W, vtx_color=vcol)
    (rc["rgb"] * torch.from_numpy(wc)).sum().backward()
    assert np.abs(vcol.grad.numpy() - gv).max() <= 1e-5 * np.abs(vcol.grad.numpy()).max()

    # public API: the texture is a parameter of the mesh, a custom loss reaches it, the optimizer changes it, the render follows
    import diffdope as dd

    m = dd.Mesh(os.path.join(su.DATA, "mesh", "AlphabetSoup.ply"), scale=su.SCALE)
    m.cuda()
    m.set_batchsize(B)
    m.enable_gradients_texture()
    assert [k for k, _ in m.named_parameters()] == ["_tex"] and m.tex.requires_grad and tuple(m.tex.shape) == (B, 2048, 2048, 3)
    proj = torch.from_numpy(ex.P.astype(np.float32)).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=0.5)

    def render():
        out = m()
        return dd.render_texture_batch(None, proj, mtx, out["pos"], out["pos_idx"], [ex.H, ex.W], uv=out["uv"], uv_idx=out["uv_idx"], tex=out["tex"])

    wt = torch.from_numpy(w).cuda()
    r0 = render()
    loss0 = (r0["rgb"] * wt).sum()
    loss0.backward()
    assert tuple(m._tex.grad.shape) == (2048, 2048, 3)
    assert np.abs(m._tex.grad.cpu().numpy() - go).max() <= 1e-5 * np.abs(go).max(), "same gradient through render_texture_batch + autograd"
    opt.step()
    r1 = render()
    loss1 = (r1["rgb"] * wt).sum()
    assert float(loss1) < float(loss0), "a gradient step on the texture lowers the loss, and the scene's copy followed the parameter"
    exp = float(loss0) - 0.5 * float((m._tex.grad ** 2).sum())  # the render is linear in the texels
    assert abs(float(loss1) - exp) <= 1e-3 * abs(float(loss0) - exp) + 1e-3 * abs(exp)
