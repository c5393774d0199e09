"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's
hot path -- `DiffDope.run_optimization` and everything it calls per iteration
(`diffdope/diffdope.py:46-89,143-234,534-613,1085-1098,1348-1375,1634-1714`).

The reference's torch glue is restated in torch (CPU, float32) so autograd gives
the same gradient algebra as the reference; the four nvdiffrast ops are the numpy
restatements in `oracle/nvdr.py` wrapped as autograd Functions with nvdiffrast's
published hand-written backward; `dd.xfm_points` follows `diffdope/c_src/mesh.cu`.

PARITY UNPINNED for the nvdiffrast ops (see `oracle/nvdr.py`). The parts that live
in the reference tree are pinned by tests/test_oracle_pins.py: SURVEY.md Appendix D, and golden
vectors produced by the reference's own functions (pose -> matrix, projection, the three losses with
their gradients, image loading; tests/golden/make_reference_vectors.py), and the output of the
reference's own unmodified run_optimization loop executed on the CPU with the four nvdiffrast ops served
by oracle/nvdr.py (tests/golden/make_reference_run.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Crop window: the reference always renders the full frame. A window (y0,x0,h,w)
is defined as "render the full frame, then slice render and ground truth"
(SURVEY.md Appendix B, row "crops"); losses average over the window's pixels.
"""
import math

import numpy as np
import torch

from . import nvdr

F = np.float32


# ----------------------------------------------------------------------------
# autograd wrappers


class _XfmPoints(torch.autograd.Function):
    """`dd.xfm_points` (`diffdope/ops.py:104-149`): fwd `mesh.cu:22-54`, bwd
    `mesh.cu:56-163` (d_points = M^T d_out, d_M = sum_n d_out (x) [p,1])."""

    @staticmethod
    def forward(ctx, points, matrix):
        ctx.save_for_backward(points, matrix)
        out = nvdr.canonical_xfm_points(points.detach().numpy(), matrix.detach().numpy())
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, dout):
        points, matrix = ctx.saved_tensors
        gp = gm = None
        ph = torch.nn.functional.pad(points, (0, 1), value=1.0)
        if ctx.needs_input_grad[1]:
            gm = torch.einsum("bnr,bnc->brc", dout, ph.expand(dout.shape[0], -1, -1))
        if ctx.needs_input_grad[0]:
            gp = torch.einsum("bnr,brc->bnc", dout, matrix)[..., :3]
        return gp, gm


def xfm_points(points, matrix):
    if points.dim() == 2:
        points = points[None]
    return _XfmPoints.apply(points, matrix)


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos_clip, tri, H, W, face_sign=None):
        rast = nvdr.rasterize(pos_clip.detach().numpy(), tri, H, W, face_sign)
        ctx.tri, ctx.H, ctx.W = tri, H, W
        r = torch.from_numpy(rast)
        ctx.save_for_backward(pos_clip, r)
        return r

    @staticmethod
    def backward(ctx, d_rast):
        pos_clip, rast = ctx.saved_tensors
        g = nvdr.rasterize_grad(pos_clip.detach().numpy(), ctx.tri, rast.numpy(), d_rast.numpy(), ctx.H, ctx.W)
        return torch.from_numpy(g), None, None, None, None


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        ctx.tri = tri
        ctx.save_for_backward(attr, rast)
        return torch.from_numpy(nvdr.interpolate(attr.detach().numpy(), rast.detach().numpy(), tri))

    @staticmethod
    def backward(ctx, d_out):
        attr, rast = ctx.saved_tensors
        d_rast, d_attr = nvdr.interpolate_grad(
            attr.detach().numpy(), rast.detach().numpy(), ctx.tri, d_out.numpy(), need_attr_grad=ctx.needs_input_grad[0]
        )
        return (torch.from_numpy(d_attr) if d_attr is not None else None), torch.from_numpy(d_rast), None


class _TextureLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tex, uv):
        ctx.save_for_backward(tex, uv)
        return torch.from_numpy(nvdr.texture_linear(tex.detach().numpy(), uv.detach().numpy()))

    @staticmethod
    def backward(ctx, d_out):
        tex, uv = ctx.saved_tensors
        g_tex = None
        if ctx.needs_input_grad[0]:  # Mesh.enable_gradients_texture (diffdope.py:909-920)
            g_tex = torch.from_numpy(nvdr.texture_linear_grad_tex(tuple(tex.shape), uv.detach().numpy(), d_out.numpy()))
        return g_tex, torch.from_numpy(nvdr.texture_linear_grad_uv(tex.detach().numpy(), uv.detach().numpy(), d_out.numpy()))


class _TextureMipmap(torch.autograd.Function):
    """EXTENSION (no reference counterpart): trilinear mip lookup, level of detail constant in the backward."""

    @staticmethod
    def forward(ctx, uv, lod, levels):
        ctx.levels = levels
        ctx.save_for_backward(uv, lod)
        return torch.from_numpy(nvdr.texture_mipmap(levels, uv.detach().numpy(), lod.numpy()))

    @staticmethod
    def backward(ctx, d_out):
        uv, lod = ctx.saved_tensors
        return torch.from_numpy(nvdr.texture_mipmap_grad_uv(ctx.levels, uv.detach().numpy(), lod.numpy(), d_out.numpy())), None, None


class _Antialias(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color, rast, pos_clip, tri, opp):
        out, work = nvdr.antialias(color.detach().numpy(), rast.detach().numpy(), pos_clip.detach().numpy(), tri, opp)
        ctx.work, ctx.tri = work, tri
        ctx.save_for_backward(color, rast, pos_clip)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, d_out):
        color, rast, pos_clip = ctx.saved_tensors
        g_color, g_pos = nvdr.antialias_grad(
            color.detach().numpy(), rast.numpy(), pos_clip.detach().numpy(), ctx.tri, ctx.work, d_out.numpy()
        )
        return torch.from_numpy(g_color), None, torch.from_numpy(g_pos), None, None


def _with_value(t, value_np):
    """Tensor with `t`'s autograd graph and exactly `value_np` as its value."""
    v = torch.from_numpy(np.ascontiguousarray(value_np, dtype=F))
    return t + (v - t).detach()


# ----------------------------------------------------------------------------
# the reference's per-iteration graph


def matrix_batch_44_from_position_quat(q, p):
    """`diffdope/diffdope.py:46-89`, same expression order."""
    r0 = torch.stack(
        [1.0 - 2.0 * q[:, 1] ** 2 - 2.0 * q[:, 2] ** 2, 2.0 * q[:, 0] * q[:, 1] - 2.0 * q[:, 2] * q[:, 3], 2.0 * q[:, 0] * q[:, 2] + 2.0 * q[:, 1] * q[:, 3]],
        dim=1,
    )
    r1 = torch.stack(
        [2.0 * q[:, 0] * q[:, 1] + 2.0 * q[:, 2] * q[:, 3], 1.0 - 2.0 * q[:, 0] ** 2 - 2.0 * q[:, 2] ** 2, 2.0 * q[:, 1] * q[:, 2] - 2.0 * q[:, 0] * q[:, 3]],
        dim=1,
    )
    r2 = torch.stack(
        [2.0 * q[:, 0] * q[:, 2] - 2.0 * q[:, 1] * q[:, 3], 2.0 * q[:, 1] * q[:, 2] + 2.0 * q[:, 0] * q[:, 3], 1.0 - 2.0 * q[:, 0] ** 2 - 2.0 * q[:, 1] ** 2],
        dim=1,
    )
    rr = torch.stack([r0, r1, r2], dim=1)
    rr = torch.cat([rr, p.reshape(-1, 3, 1)], dim=2)
    bottom = torch.tensor([0, 0, 0, 1], dtype=torch.float32).expand(p.shape[0], 1, 4)
    return torch.cat([rr, bottom], dim=1)


class Mesh:
    """Host arrays of one object, loaded the way `Mesh.__init__` does
    (`diffdope/diffdope.py:784-851`): pos*scale, int32 faces, v -> 1-v, tex/255."""

    def __init__(self, pos, tri, uv=None, tex=None, vtx_color=None):
        self.pos = np.ascontiguousarray(pos, dtype=F)
        self.tri = np.ascontiguousarray(tri, dtype=np.int64)
        self.uv = None if uv is None else np.ascontiguousarray(uv, dtype=F)
        self.tex = None if tex is None else np.ascontiguousarray(tex, dtype=F)
        self.vtx_color = None if vtx_color is None else np.ascontiguousarray(vtx_color, dtype=F)
        self.opp = nvdr.build_edge_opposites(self.tri)
        self.cull_sign = nvdr.closed_mesh_orientation(self.pos, self.tri)  # 0: open / inconsistent mesh, nothing is culled
        self.cull = True  # back-face culling of closed meshes (the raster rule); False = rasterise every triangle
        self.texture_filter = "linear"  # or "linear-mipmap-linear" (extension)
        self._mips = None

    def mip_levels(self):
        if self._mips is None:
            self._mips = nvdr.build_mip_chain(self.tex)
        return self._mips

    @property
    def textured(self):
        return self.tex is not None


def render(mesh, proj, quat_raw, trans, H, W, tex=None, vtx_color=None):
    """`Object3D.forward` + `matrix_batch_44_from_position_quat` + `render_texture_batch`
    (`diffdope/diffdope.py:1085-1098,46-89,156-234`). quat_raw/trans are torch leaf
    (or any) tensors [B,4]/[B,3]. Returns dict rgb [B,H,W,3], depth [B,H,W],
    mask [B,H,W,3], rast_out, mtx. `tex` / `vtx_color`: torch tensors to use instead of the mesh's arrays (so that they can
    require grad: `Mesh.enable_gradients_texture`, diffdope.py:909-920)."""
    B = quat_raw.shape[0]
    q = quat_raw / torch.norm(quat_raw, dim=1).reshape(-1, 1)
    mtx = matrix_batch_44_from_position_quat(q, trans)
    qh_c, M_c = nvdr.canonical_pose(quat_raw.detach().numpy(), trans.detach().numpy())
    mtx = _with_value(mtx, M_c)
    proj_t = torch.from_numpy(np.asarray(proj, dtype=F))
    mvp = torch.matmul(proj_t.expand(B, 4, 4), mtx)
    mvp = _with_value(mvp, nvdr.canonical_mvp(proj, M_c))

    pos = torch.from_numpy(mesh.pos)
    pos_b = pos[None].expand(B, -1, -1)
    pos_clip = xfm_points(pos_b, mvp)
    face = nvdr.face_signs(mesh.cull_sign if mesh.cull else 0, proj, M_c, mesh.pos.min(0), mesh.pos.max(0))
    rast = _Rasterize.apply(pos_clip, mesh.tri, H, W, face)

    posw = torch.cat([pos, torch.ones(pos.shape[0], 1)], dim=1)
    gb_pos = _Interpolate.apply(posw, rast, mesh.tri)
    depth = xfm_points(gb_pos.reshape(B, -1, 4)[..., :3], mtx)
    depth = depth.reshape(B, H, W, 4)[..., 2] * -1

    ones = torch.ones(pos.shape[0], 3)
    mask = _Interpolate.apply(ones, rast, mesh.tri)
    mask = _Antialias.apply(mask, rast, pos_clip, mesh.tri, mesh.opp)

    if mesh.textured:
        texc = _Interpolate.apply(torch.from_numpy(mesh.uv), rast, mesh.tri)
        if mesh.texture_filter == "linear-mipmap-linear":
            levels = mesh.mip_levels()
            lod = nvdr.texture_lod(pos_clip.detach().numpy(), mesh.tri, mesh.uv, rast.detach().numpy(), mesh.tex.shape[:2], len(levels))
            color = _TextureMipmap.apply(texc, torch.from_numpy(lod), levels)
        else:
            color = _TextureLinear.apply(torch.from_numpy(mesh.tex) if tex is None else tex, texc)
    else:
        color = _Interpolate.apply(torch.from_numpy(mesh.vtx_color) if vtx_color is None else vtx_color, rast, mesh.tri)
    color = color * torch.clamp(rast[..., -1:], 0, 1)
    return {"rgb": color, "depth": depth, "mask": mask, "rast_out": rast, "mtx": mtx}


_SOBEL = torch.tensor([[[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], [[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]]])


def sobel_magnitude(img):
    """EXTENSION: [B,h,w,3] -> [B,h,w]; grey = channel mean, 3x3 Sobel with zero
    padding (cross-correlation, row index increasing with the array's first image axis),
    sqrt(Gx^2 + Gy^2 + 1e-12)."""
    g = img.mean(dim=-1).unsqueeze(1)
    d = torch.nn.functional.conv2d(g, _SOBEL.unsqueeze(1), padding=1)
    return torch.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + 1e-12)


def window_slice(t, window):
    if window is None:
        return t
    y0, x0, h, w = window
    return t[:, y0 : y0 + h, x0 : x0 + w]


def losses(renders, gt, lr_mult, cfg_losses, window=None):
    """`l1_rgb_with_mask`, `l1_depth_with_mask`, `l1_mask` with `dist_batch_lr`
    (`diffdope/diffdope.py:534-613`). gt: dict of unbatched tensors rgb [H,W,3],
    depth [H,W], segmentation [H,W,3]. Returns (loss scalar, {key: [B] logged value})."""
    total = torch.zeros(1)
    logged = {}
    seg = window_slice(gt["segmentation"][None], window)
    if cfg_losses.get("l1_rgb_with_mask"):
        w = cfg_losses.get("weight_rgb", 1.0)
        d = torch.abs((window_slice(renders["rgb"], window) - window_slice(gt["rgb"][None], window)) * seg)
        logged["rgb"] = torch.mean(d.detach(), (1, 2, 3)) * w
        total = total + (torch.mean(d, (1, 2, 3)) * lr_mult).mean() * w
    if cfg_losses.get("l1_depth_with_mask"):
        w = cfg_losses.get("weight_depth", 1.0)
        d = torch.abs((window_slice(renders["depth"], window) - window_slice(gt["depth"][None], window)) * seg[..., 0])
        logged["depth"] = torch.mean(d.detach(), (1, 2)) * w
        total = total + (torch.mean(d, (1, 2)) * lr_mult).mean() * w
    if cfg_losses.get("l1_mask"):
        w = cfg_losses.get("weight_mask", 1.0)
        d = torch.abs(window_slice(renders["mask"], window) - seg)
        logged["mask_selection"] = torch.mean(torch.abs(d.detach()), (1, 2, 3)) * w
        total = total + (torch.mean(d, (1, 2, 3)) * lr_mult).mean() * w
    if cfg_losses.get("l1_edge"):  # extension, no reference counterpart
        w = cfg_losses.get("weight_edge", 1.0)
        er = sobel_magnitude(window_slice(renders["rgb"], window))
        eg = sobel_magnitude(window_slice(gt["rgb"][None], window))
        d = torch.abs((er - eg) * seg[..., 0])
        logged["edge"] = torch.mean(d.detach(), (1, 2)) * w
        total = total + (torch.mean(d, (1, 2)) * lr_mult).mean() * w
    return total, logged


def lr_schedule(it, nb_iterations, base_lr, lr_decay):
    """`diffdope/diffdope.py:1657-1661`."""
    itf = it / nb_iterations + 1
    return base_lr * lr_decay**itf


def forward_backward(mesh, proj, quat_raw, trans, gt, lr_mult, cfg_losses, H, W, window=None, b_global=None):
    """One forward + backward. Returns (logged losses, d_quat [B,4], d_trans [B,3], renders).
    `b_global`: divisor of the hypothesis mean when the batch is a shard."""
    q = torch.tensor(np.asarray(quat_raw, dtype=F), requires_grad=True)
    t = torch.tensor(np.asarray(trans, dtype=F), requires_grad=True)
    lr = torch.as_tensor(np.asarray(lr_mult, dtype=F))
    r = render(mesh, proj, q, t, H, W)
    total, logged = losses(r, gt, lr, cfg_losses, window)
    if b_global is not None:
        total = total * (q.shape[0] / float(b_global))
    total.backward()
    return logged, q.grad.numpy().copy(), t.grad.numpy().copy(), r


def run_optimization(mesh, proj, quat0, trans0, gt, lr_mult, cfg_losses, hyper, H, W, window=None, progress=False, b_global=None, stop_after=None):
    """`DiffDope.run_optimization` (`diffdope/diffdope.py:1634-1714`): nb_iterations+1
    iterations of forward, logging, loss, backward, SGD step with the decayed rate.

    quat0 [B,4], trans0 [B,3]: initial parameter values (the reference starts every
    hypothesis at the same pose, `diffdope.py:1019-1026`). Returns dict with
    `poses` [iters, B, 7] (the parameters each iteration rendered with),
    `mtx` [iters, B, 4, 4], `losses` {key: [iters, B]}, `final` (q, t after the last step).
    `b_global`: divisor of the hypothesis mean when the batch is a shard of a larger job.
    `stop_after`: run only the first `stop_after` iterations of the schedule (tests at full size)."""
    nb = int(hyper["nb_iterations"])
    params = [torch.nn.Parameter(torch.tensor(np.asarray(quat0, dtype=F)[:, i].copy())) for i in range(4)]
    params += [torch.nn.Parameter(torch.tensor(np.asarray(trans0, dtype=F)[:, i].copy())) for i in range(3)]
    if hyper.get("optimizer", "sgd") == "adam":  # extension: torch.optim.Adam on the same graph
        opt = torch.optim.Adam(params, lr=hyper.get("learning_rate_base", 1), betas=(hyper.get("adam_beta1", 0.9), hyper.get("adam_beta2", 0.999)),
                               eps=hyper.get("adam_eps", 1e-8))
    else:
        opt = torch.optim.SGD(params, lr=hyper.get("learning_rate_base", 1))
    lr = torch.as_tensor(np.asarray(lr_mult, dtype=F))
    poses, mtxs, hist = [], [], {}
    for it in range(nb + 1 if stop_after is None else min(nb + 1, int(stop_after))):
        lr_t = lr_schedule(it, nb, hyper["base_lr"], hyper["lr_decay"])
        for g in opt.param_groups:
            g["lr"] = lr_t
        opt.zero_grad()
        q = torch.stack(params[:4], dim=0).T
        t = torch.stack(params[4:], dim=0).T
        poses.append(torch.cat([q, t], dim=1).detach().numpy().copy())
        r = render(mesh, proj, q, t, H, W)
        mtxs.append(r["mtx"].detach().numpy().copy())
        total, logged = losses(r, gt, lr, cfg_losses, window)
        for k, v in logged.items():
            hist.setdefault(k, []).append(v.numpy().copy())
        if b_global is not None:
            total = total * (q.shape[0] / float(b_global))
        total.backward()
        opt.step()
        if progress:
            print("it %d loss %.6f" % (it, float(total)))
    final = np.stack([p.detach().numpy() for p in params], axis=1)
    return {
        "poses": np.stack(poses),
        "mtx": np.stack(mtxs),
        "losses": {k: np.stack(v) for k, v in hist.items()},
        "final": final,
    }


def argmin_hypothesis(loss_hist):
    """`DiffDope.get_argmin` (`diffdope/diffdope.py:1488-1513`)."""
    last = np.stack([v[-1] for v in loss_hist.values()], axis=0)
    return int(np.argmin(last.mean(axis=0)))


# ----------------------------------------------------------------------------
# scene setup helpers shared by tests / bench (host side, restating the reference's loaders)


def projection_matrix(fx, fy, cx, cy, w, h, znear=0.01, zfar=200.0):
    """`Camera.get_projection_matrix`, y_down branch (`diffdope/diffdope.py:679-742`)."""
    depth = float(zfar - znear)
    q = -(zfar + znear) / depth
    qn = -2 * (zfar * znear) / depth
    return np.array(
        [[2 * fx / w, 0.0, (-2 * cx + w) / w, 0], [0, 2 * fy / h, (2 * cy - h) / h, 0], [0, 0, q, qn], [0, 0, -1, 0]],
        dtype=np.float64,
    )
