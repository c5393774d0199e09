"""ORACLE (test infrastructure, never shipped): numpy float32 restatement of the
four nvdiffrast ops the reference's hot path calls, forward and backward.

PARITY UNPINNED. The reference gets this arithmetic from `nvdiffrast.torch`
(`diffdope/diffdope.py:25,147,198,212-214,218-226,230`), an un-vendored and
un-pinned git dependency (`setup.py:14`, `requirements.txt:7`) that is not on
disk here. The functions below restate nvdiffrast's *published* algorithm
(Laine et al. 2020 and the public v0.3.x CUDA sources: `RasterizeCudaFwdShaderKernel`,
`RasterizeGradKernel`, `InterpolateFwd/GradKernel`, `TextureFwd/GradKernelLinear`,
`AntialiasFwdAnalysisKernel`, `AntialiasGradKernel`), anchored on the reference's
call sites. No golden vector of the reference exists for any of them
(SURVEY.md section 4).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Coverage contract (shared bit-for-bit with the CUDA kernels, DESIGN.md "raster rule"):
the reference rasterises with the OpenGL hardware rasteriser whose sub-pixel
snapping is implementation defined, so the rule is fixed here instead:
pixel-centre sampling, window coordinates snapped to 1/256 px (round half to
even), 64-bit integer edge functions, inward-normal tie rule, depth test LESS on
the shader's float32 z/w with the lower triangle index winning ties, fragments
with z/w outside [-1,1] discarded, triangles with any w <= 0 culled (no
near-plane clipping).

Every float32 operation on the decision path (pose -> matrix -> clip -> snap ->
z/w -> barycentrics -> antialias analysis) is a separately rounded IEEE
add/sub/mul/div/sqrt in the order written, which the CUDA side mirrors with
__fmul_rn/__fadd_rn/__fdiv_rn, so discrete decisions agree exactly.
"""
import numpy as np

F = np.float32
SUBPIX = 256
FLT_MAX = np.finfo(np.float32).max
EMPTY_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)
COORD_LIMIT = float(1 << 20)


# ----------------------------------------------------------------------------
# canonical pose -> matrices -> clip space


def canonical_pose(quat_raw, trans):
    """(q/|q|, M[B,4,4]) from raw quaternion params [B,4] (x,y,z,w) and translation [B,3].

    Same algebra as `Object3D.forward` (`diffdope/diffdope.py:1090-1096`) followed by
    `matrix_batch_44_from_position_quat` (`:46-89`), in a fixed float32 operation order.
    """
    q = np.asarray(quat_raw, dtype=F)
    t = np.asarray(trans, dtype=F)
    n = np.sqrt(((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]) + q[:, 3] * q[:, 3])
    qh = q / n[:, None]
    q0, q1, q2, q3 = qh[:, 0], qh[:, 1], qh[:, 2], qh[:, 3]
    one, two = F(1.0), F(2.0)
    B = q.shape[0]
    M = np.zeros((B, 4, 4), dtype=F)
    M[:, 0, 0] = (one - two * (q1 * q1)) - two * (q2 * q2)
    M[:, 0, 1] = (two * q0) * q1 - (two * q2) * q3
    M[:, 0, 2] = (two * q0) * q2 + (two * q1) * q3
    M[:, 1, 0] = (two * q0) * q1 + (two * q2) * q3
    M[:, 1, 1] = (one - two * (q0 * q0)) - two * (q2 * q2)
    M[:, 1, 2] = (two * q1) * q2 - (two * q0) * q3
    M[:, 2, 0] = (two * q0) * q2 - (two * q1) * q3
    M[:, 2, 1] = (two * q1) * q2 + (two * q0) * q3
    M[:, 2, 2] = (one - two * (q0 * q0)) - two * (q1 * q1)
    M[:, 0, 3] = t[:, 0]
    M[:, 1, 3] = t[:, 1]
    M[:, 2, 3] = t[:, 2]
    M[:, 3, 3] = one
    return qh, M


def canonical_mvp(proj, M):
    """MVP = P @ M (`diffdope/diffdope.py:195`), fixed float32 order."""
    P = np.asarray(proj, dtype=F)
    M = np.asarray(M, dtype=F)
    B = M.shape[0]
    out = np.zeros((B, 4, 4), dtype=F)
    for r in range(4):
        for c in range(4):
            out[:, r, c] = ((P[r, 0] * M[:, 0, c] + P[r, 1] * M[:, 1, c]) + P[r, 2] * M[:, 2, c]) + P[r, 3] * M[:, 3, c]
    return out


def canonical_xfm_points(pos, mtx):
    """out[b,v,:] = M[b] @ [p_v, 1] (`diffdope/c_src/mesh.cu:22-54`, `ops.py:137-141`)."""
    p = np.asarray(pos, dtype=F)
    if p.ndim == 3:
        x, y, z = p[:, :, 0], p[:, :, 1], p[:, :, 2]
    else:
        x, y, z = p[None, :, 0], p[None, :, 1], p[None, :, 2]
    M = np.asarray(mtx, dtype=F)
    cols = []
    for r in range(4):
        m0, m1, m2, m3 = (M[:, r, k][:, None] for k in range(4))
        cols.append(((m0 * x + m1 * y) + m2 * z) + m3)
    return np.stack(cols, axis=-1)


# ----------------------------------------------------------------------------
# topology: opposite vertex across each triangle edge


def build_edge_opposites(tri):
    """opp[t, i] = the vertex opposite edge i of triangle t in the *other* triangle
    that shares that edge (by vertex index), or -1.

    Edge i of a triangle (v0,v1,v2) is the edge not containing v_i. Deterministic
    restatement of nvdiffrast's edge-vertex hash (`AntialiasFwdMeshKernel`): an
    edge keeps the first two opposite vertices inserted, in triangle order;
    a lookup returns the stored vertex that differs from the caller's own.
    """
    tri = np.asarray(tri, dtype=np.int64)
    T = tri.shape[0]
    slots = {}
    for t in range(T):
        v = tri[t]
        for i in range(3):
            a, b, c = int(v[(i + 1) % 3]), int(v[(i + 2) % 3]), int(v[i])
            key = (a, b) if a < b else (b, a)
            s = slots.get(key)
            if s is None:
                slots[key] = [c, -1]
            elif s[1] < 0:
                s[1] = c
    opp = np.full((T, 3), -1, dtype=np.int32)
    for t in range(T):
        v = tri[t]
        for i in range(3):
            a, b, c = int(v[(i + 1) % 3]), int(v[(i + 2) % 3]), int(v[i])
            key = (a, b) if a < b else (b, a)
            s = slots[key]
            o = s[0] if s[0] != c else s[1]
            opp[t, i] = o
    return opp


# ----------------------------------------------------------------------------
# rasterize


def _float_key(zw):
    """Order-preserving uint32 image of a float32 (zw in [-1,1])."""
    bits = zw.view(np.uint32)
    neg = (bits >> 31).astype(bool)
    return np.where(neg, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)


def snap_vertices(clip, W, H):
    """Window coordinates in 1/256 px, int64 [B,V], plus per-vertex validity."""
    clip = np.asarray(clip, dtype=F)
    x, y, w = clip[..., 0], clip[..., 1], clip[..., 3]
    hw, hh = F(W) * F(0.5), F(H) * F(0.5)
    with np.errstate(all="ignore"):
        sx = (x / w) * hw + hw
        sy = (y / w) * hh + hh
        ok = (w > 0) & np.isfinite(sx) & np.isfinite(sy) & (np.abs(sx) < COORD_LIMIT) & (np.abs(sy) < COORD_LIMIT)
        X = np.rint(np.where(ok, sx, 0) * F(SUBPIX)).astype(np.int64)
        Y = np.rint(np.where(ok, sy, 0) * F(SUBPIX)).astype(np.int64)
    return X, Y, ok


def _edge_fns(X, Y, tri):
    """Orientation-normalised integer edge setup. Returns per-triangle arrays."""
    x0, x1, x2 = X[tri[:, 0]], X[tri[:, 1]], X[tri[:, 2]]
    y0, y1, y2 = Y[tri[:, 0]], Y[tri[:, 1]], Y[tri[:, 2]]
    area2 = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0)
    flip = area2 < 0
    # swap v1 <-> v2 for negative orientation (coverage only)
    ax1 = np.where(flip, x2, x1)
    ay1 = np.where(flip, y2, y1)
    ax2 = np.where(flip, x1, x2)
    ay2 = np.where(flip, y1, y2)
    return (x0, y0, ax1, ay1, ax2, ay2), area2


def _inside(ev, cx, cy):
    x0, y0, x1, y1, x2, y2 = ev
    ins = np.ones(cx.shape, dtype=bool)
    for (ax, ay, bx, by) in ((x0, y0, x1, y1), (x1, y1, x2, y2), (x2, y2, x0, y0)):
        dx, dy = bx - ax, by - ay
        e = dx * (cy - ay) - dy * (cx - ax)
        nx, ny = -dy, dx  # inward normal of the orientation-normalised edge
        own = (nx > 0) | ((nx == 0) & (ny > 0))
        ins &= (e > 0) | ((e == 0) & own)
    return ins


def shader_terms(c0, c1, c2, fx, fy):
    """Homogeneous edge functions at pixel-centre NDC (fx,fy); returns a0,a1,a2."""
    p0x = c0[..., 0] - fx * c0[..., 3]
    p0y = c0[..., 1] - fy * c0[..., 3]
    p1x = c1[..., 0] - fx * c1[..., 3]
    p1y = c1[..., 1] - fy * c1[..., 3]
    p2x = c2[..., 0] - fx * c2[..., 3]
    p2y = c2[..., 1] - fy * c2[..., 3]
    a0 = p1x * p2y - p1y * p2x
    a1 = p2x * p0y - p2y * p0x
    a2 = p0x * p1y - p0y * p1x
    return a0, a1, a2, (p0x, p0y, p1x, p1y, p2x, p2y)


def pixel_ndc(px, py, W, H):
    xs, xo = F(2.0) / F(W), F(1.0) / F(W) - F(1.0)
    ys, yo = F(2.0) / F(H), F(1.0) / F(H) - F(1.0)
    return xs * px.astype(F) + xo, ys * py.astype(F) + yo


def closed_mesh_orientation(pos, tri):
    """+1 / -1 if the mesh, after welding vertices with bit-identical positions (uv seams duplicate them), is a
    closed, consistently oriented 2-manifold -- every directed edge occurs exactly once and so does its reverse --
    with positive / negative enclosed volume; 0 otherwise (no culling possible)."""
    pos = np.ascontiguousarray(pos, dtype=F)
    tri = np.asarray(tri, dtype=np.int64)
    _, weld = np.unique(pos.view(np.uint32).reshape(-1, 3), axis=0, return_inverse=True)
    t = weld.reshape(-1)[tri]
    if ((t[:, 0] == t[:, 1]) | (t[:, 1] == t[:, 2]) | (t[:, 2] == t[:, 0])).any():
        return 0
    a = np.concatenate([t[:, 0], t[:, 1], t[:, 2]])
    b = np.concatenate([t[:, 1], t[:, 2], t[:, 0]])
    n = int(weld.max()) + 1
    fwd = a * n + b
    if np.unique(fwd).size != fwd.size:
        return 0
    if not np.array_equal(np.sort(fwd), np.sort(b * n + a)):
        return 0
    p = pos.astype(np.float64)
    vol = np.einsum("ij,ij->i", p[tri[:, 0]], np.cross(p[tri[:, 1]], p[tri[:, 2]])).sum()
    return int(np.sign(vol))


def face_signs(cull_sign, proj, M, bbmin=None, bbmax=None):
    """Sign of the snapped window-space area of a FRONT-facing triangle per batch entry: mesh orientation x
    orientation of the camera-to-window map x orientation of the model matrix (0 = do not cull). A camera inside
    the object's bounding box (bbmin / bbmax, object space) sees back faces: 0 for that entry."""
    P = np.asarray(proj, dtype=np.float64)
    M = np.asarray(M, dtype=np.float64)
    dp = np.sign(P[0, 0] * P[1, 1] - P[0, 1] * P[1, 0])
    dm = np.sign(np.linalg.det(M[:, :3, :3]))
    face = (cull_sign * dp * dm).astype(np.int64)
    if bbmin is not None:
        for b in range(M.shape[0]):
            if face[b] != 0:
                cam = -np.linalg.solve(M[b, :3, :3], M[b, :3, 3])
                outside = (cam < np.asarray(bbmin, np.float64)).any() or (cam > np.asarray(bbmax, np.float64)).any()
                if not outside:
                    face[b] = 0
    return face


def _rasterize_homogeneous(clipb, tri, t, H, W, face):
    """Fragments of ONE triangle that cannot be snapped -- a vertex behind the camera plane (w <= 0) or outside the fixed-point
    range -- by homogeneous rasterisation (no clipping; the raster rule): at every pixel centre of a
    conservative bounding box the fragment formula's edge functions decide coverage (inside iff none has the opposite sign of
    their sum), the interpolated w must be positive (in front of the camera) and z/w in [-1, 1] (between near and far: what
    clipping against the near plane achieves in GL / nvdiffrast). Same float32 arithmetic as csrc/raster_common.cuh."""
    c = [clipb[tri[t, k]].astype(F) for k in range(3)]
    hw, hh = F(0.5) * F(W), F(0.5) * F(H)
    pts = []
    bad = False
    with np.errstate(all="ignore"):
        for i in range(3):
            a, b_ = c[i], c[(i + 1) % 3]
            da, db = a[2] + a[3], b_[2] + b_[3]
            if da > 0 and a[3] > 0:
                pts.append((a[0] / a[3] * hw + hw, a[1] / a[3] * hh + hh))
            if (da > 0) != (db > 0):
                tt = da / (da - db)
                x, y, w = a[0] + tt * (b_[0] - a[0]), a[1] + tt * (b_[1] - a[1]), a[3] + tt * (b_[3] - a[3])
                if w > 0:
                    pts.append((x / w * hw + hw, y / w * hh + hh))
                else:
                    bad = True
    if not pts:
        return None
    P = np.array(pts, dtype=np.float64)
    if bad or not np.all(np.isfinite(P)) or np.abs(P).max() >= 1e9:
        x0, x1, y0, y1 = 0, W - 1, 0, H - 1
    else:
        x0, x1 = max(int(np.floor(P[:, 0].min())) - 3, 0), min(int(np.ceil(P[:, 0].max())) + 2, W - 1)
        y0, y1 = max(int(np.floor(P[:, 1].min())) - 3, 0), min(int(np.ceil(P[:, 1].max())) + 2, H - 1)
    if x0 > x1 or y0 > y1:
        return None
    cpy, cpx = np.mgrid[y0:y1 + 1, x0:x1 + 1]
    cpx, cpy = cpx.reshape(-1), cpy.reshape(-1)
    fx, fy = pixel_ndc(cpx, cpy, W, H)
    c0, c1, c2 = (np.broadcast_to(v, (cpx.size, 4)) for v in c)
    with np.errstate(all="ignore"):
        a0, a1, a2, _ = shader_terms(c0, c1, c2, fx, fy)
        sm = (a0 + a1) + a2
        pos = (sm > 0) & (a0 >= 0) & (a1 >= 0) & (a2 >= 0)
        neg = (sm < 0) & (a0 <= 0) & (a1 <= 0) & (a2 <= 0)
        ins = pos | neg
        if face != 0:
            ins &= pos == (face > 0)
        z = (c0[:, 2] * a0 + c1[:, 2] * a1) + c2[:, 2] * a2
        w = (c0[:, 3] * a0 + c1[:, 3] * a1) + c2[:, 3] * a2
        ins &= ((w > 0) == pos) & (w != 0)
        zw = z / w
        ins &= (zw >= F(-1.0)) & (zw <= F(1.0))
    if not ins.any():
        return None
    return (cpx[ins], cpy[ins], np.full(int(ins.sum()), t, dtype=np.int64), zw[ins].astype(F), a0[ins], a1[ins], a2[ins])


def rasterize(clip, tri, H, W, face_sign=None):
    """Restates `dr.rasterize(glctx, pos_clip, tri, [H,W])` (`diffdope/diffdope.py:198-200`).

    face_sign [B] (optional, from `face_signs`): where non-zero, triangles whose snapped area has the other sign
    are back faces of a closed mesh and are skipped. They can never be the front-most surface, so coverage is
    unchanged; the winner can differ from a no-culling rasteriser only where a back and a front face tie in depth
    within float rounding on a silhouette (the raster rule).

    Triangles with a vertex behind the camera plane (w <= 0) or outside the fixed-point range are not dropped (GL / nvdiffrast clip
    them): see `_rasterize_homogeneous`. Triangles entirely behind the camera plane are.

    clip [B,V,4] float32, tri [T,3] int. Returns rast_out [B,H,W,4] float32 =
    (u, v, z/w, tri_id+1), all-zero at background. The pixel-derivative output
    `rast_db` is not produced: every consumer of it in the reference is discarded
    (`diffdope.py:203,212-213,218-226` with filter_mode="linear").
    """
    clip = np.ascontiguousarray(clip, dtype=F)
    tri = np.asarray(tri, dtype=np.int64)
    B = clip.shape[0]
    T = tri.shape[0]
    rast = np.zeros((B, H, W, 4), dtype=F)
    for b in range(B):
        X, Y, ok = snap_vertices(clip[b], W, H)
        ev, area2 = _edge_fns(X, Y, tri)
        tri_ok = ok[tri[:, 0]] & ok[tri[:, 1]] & ok[tri[:, 2]] & (area2 != 0)
        if face_sign is not None and face_sign[b] != 0:
            tri_ok &= np.sign(area2) == face_sign[b]
        xmin = np.minimum(np.minimum(ev[0], ev[2]), ev[4])
        xmax = np.maximum(np.maximum(ev[0], ev[2]), ev[4])
        ymin = np.minimum(np.minimum(ev[1], ev[3]), ev[5])
        ymax = np.maximum(np.maximum(ev[1], ev[3]), ev[5])
        half = SUBPIX // 2
        pxmin = np.maximum((xmin - half + SUBPIX - 1) >> 8, 0)
        pxmax = np.minimum((xmax - half) >> 8, W - 1)
        pymin = np.maximum((ymin - half + SUBPIX - 1) >> 8, 0)
        pymax = np.minimum((ymax - half) >> 8, H - 1)
        tri_ok &= (pxmin <= pxmax) & (pymin <= pymax)
        ids = np.nonzero(tri_ok)[0]
        cand = []  # (cpx, cpy, ct, zw, a0, a1, a2) of every fragment that passed coverage and the depth range
        if ids.size > 0:
            nx = (pxmax - pxmin + 1)[ids]
            ny = (pymax - pymin + 1)[ids]
            # candidate (triangle, pixel) pairs: enumerate the bbox of every triangle
            counts = nx * ny
            total = int(counts.sum())
            rep = np.repeat(np.arange(ids.size), counts)
            offs = np.arange(total) - np.repeat(np.cumsum(counts) - counts, counts)
            cpx = pxmin[ids][rep] + offs % nx[rep]
            cpy = pymin[ids][rep] + offs // nx[rep]
            ct = ids[rep]
            evc = tuple(e[ct] for e in ev)
            ins = _inside(evc, cpx * SUBPIX + half, cpy * SUBPIX + half)
            cpx, cpy, ct = cpx[ins], cpy[ins], ct[ins]
            if ct.size > 0:
                c0, c1, c2 = clip[b][tri[ct, 0]], clip[b][tri[ct, 1]], clip[b][tri[ct, 2]]
                fx, fy = pixel_ndc(cpx, cpy, W, H)
                with np.errstate(all="ignore"):
                    a0, a1, a2, _ = shader_terms(c0, c1, c2, fx, fy)
                    z = (c0[:, 2] * a0 + c1[:, 2] * a1) + c2[:, 2] * a2
                    w = (c0[:, 3] * a0 + c1[:, 3] * a1) + c2[:, 3] * a2
                    zw = z / w
                    keep = (zw >= F(-1.0)) & (zw <= F(1.0))  # NaN fails both
                cand.append((cpx[keep], cpy[keep], ct[keep], zw[keep], a0[keep], a1[keep], a2[keep]))
        # triangles that cross the camera plane (a vertex with w <= 0) or leave the fixed-point range: homogeneous rasterisation
        wv = clip[b][:, 3]
        all_ok = ok[tri[:, 0]] & ok[tri[:, 1]] & ok[tri[:, 2]]
        some_front = (wv[tri[:, 0]] > 0) | (wv[tri[:, 1]] > 0) | (wv[tri[:, 2]] > 0)
        for t in np.nonzero(~all_ok & some_front)[0]:
            c = _rasterize_homogeneous(clip[b], tri, int(t), H, W, 0 if face_sign is None else int(face_sign[b]))
            if c is not None:
                cand.append(c)
        if not cand:
            continue
        cpx, cpy, ct, zw, a0, a1, a2 = (np.concatenate([c[k] for c in cand]) for k in range(7))
        if ct.size == 0:
            continue
        key = (_float_key(zw).astype(np.uint64) << np.uint64(32)) | ct.astype(np.uint64)
        pix = cpy * W + cpx
        zbuf = np.full(H * W, EMPTY_KEY, dtype=np.uint64)
        np.minimum.at(zbuf, pix, key)
        win = zbuf[pix] == key
        with np.errstate(all="ignore"):
            iw = F(1.0) / ((a0[win] + a1[win]) + a2[win])
            u = np.clip(a0[win] * iw, F(0.0), F(1.0))
            v = np.clip(a1[win] * iw, F(0.0), F(1.0))
        u = np.where(np.isnan(u), F(0.0), u)
        v = np.where(np.isnan(v), F(0.0), v)
        r = rast[b].reshape(H * W, 4)
        r[pix[win], 0] = u
        r[pix[win], 1] = v
        r[pix[win], 2] = zw[win]
        r[pix[win], 3] = (ct[win] + 1).astype(F)
    return rast


def rasterize_grad(clip, tri, rast, d_rast, H, W):
    """nvdiffrast `RasterizeGradKernel`: gradient of (u,v) w.r.t. clip (x,y,w) with the
    triangle id held fixed; z/w and id are non-differentiable; clamps are ignored;
    the denominator carries copysign(1e-6, sum)."""
    clip = np.asarray(clip, dtype=F)
    B, V = clip.shape[0], clip.shape[1]
    g = np.zeros((B, V, 4), dtype=F)
    tid = rast[..., 3].astype(np.int64) - 1
    b, py, px = np.nonzero(tid >= 0)
    if b.size == 0:
        return g
    t = tid[b, py, px]
    dy0 = d_rast[b, py, px, 0].astype(F)
    dy1 = d_rast[b, py, px, 1].astype(F)
    nz = (dy0 != 0) | (dy1 != 0)
    b, py, px, t, dy0, dy1 = b[nz], py[nz], px[nz], t[nz], dy0[nz], dy1[nz]
    v0, v1, v2 = tri[t, 0], tri[t, 1], tri[t, 2]
    c0, c1, c2 = clip[b, v0], clip[b, v1], clip[b, v2]
    fx, fy = pixel_ndc(px, py, W, H)
    a0, a1, a2, (p0x, p0y, p1x, p1y, p2x, p2y) = shader_terms(c0, c1, c2, fx, fy)
    at = (a0 + a1) + a2
    ep = np.copysign(F(1e-6), at)
    iw = F(1.0) / (at + ep)
    b0, b1 = a0 * iw, a1 * iw
    gb0, gb1 = dy0 * iw, dy1 * iw
    gbb = gb0 * b0 + gb1 * b1
    gp0x = gbb * (p2y - p1y) - gb1 * p2y
    gp1x = gbb * (p0y - p2y) + gb0 * p2y
    gp2x = gbb * (p1y - p0y) - gb0 * p1y + gb1 * p0y
    gp0y = gbb * (p1x - p2x) + gb1 * p2x
    gp1y = gbb * (p2x - p0x) - gb0 * p2x
    gp2y = gbb * (p0x - p1x) + gb0 * p1x - gb1 * p0x
    gp0w = -fx * gp0x - fy * gp0y
    gp1w = -fx * gp1x - fy * gp1y
    gp2w = -fx * gp2x - fy * gp2y
    zero = np.zeros_like(gp0x)
    np.add.at(g, (b, v0), np.stack([gp0x, gp0y, zero, gp0w], -1))
    np.add.at(g, (b, v1), np.stack([gp1x, gp1y, zero, gp1w], -1))
    np.add.at(g, (b, v2), np.stack([gp2x, gp2y, zero, gp2w], -1))
    return g


# ----------------------------------------------------------------------------
# interpolate


def _attr_at(attr, b, idx):
    return attr[b, idx] if attr.ndim == 3 else attr[idx]


def interpolate(attr, rast, tri):
    """`dr.interpolate(attr, rast, tri)` main output: u*a0 + v*a1 + (1-u-v)*a2, zero at
    background. attr [V,A] or [B,V,A]."""
    attr = np.asarray(attr, dtype=F)
    B, H, W, _ = rast.shape
    A = attr.shape[-1]
    out = np.zeros((B, H, W, A), dtype=F)
    tid = rast[..., 3].astype(np.int64) - 1
    b, py, px = np.nonzero(tid >= 0)
    t = tid[b, py, px]
    b0 = rast[b, py, px, 0][:, None]
    b1 = rast[b, py, px, 1][:, None]
    b2 = (F(1.0) - b0) - b1
    a0 = _attr_at(attr, b, tri[t, 0])
    a1 = _attr_at(attr, b, tri[t, 1])
    a2 = _attr_at(attr, b, tri[t, 2])
    out[b, py, px] = (b0 * a0 + b1 * a1) + b2 * a2
    return out


def interpolate_grad(attr, rast, tri, d_out, need_attr_grad=False):
    """`InterpolateGradKernel`: d_rast (u,v) = sum_c dy*(a0-a2), sum_c dy*(a1-a2);
    d_attr by scatter (only when asked for)."""
    attr = np.asarray(attr, dtype=F)
    d_rast = np.zeros(rast.shape, dtype=F)
    tid = rast[..., 3].astype(np.int64) - 1
    b, py, px = np.nonzero(tid >= 0)
    t = tid[b, py, px]
    a0 = _attr_at(attr, b, tri[t, 0])
    a1 = _attr_at(attr, b, tri[t, 1])
    a2 = _attr_at(attr, b, tri[t, 2])
    dy = d_out[b, py, px].astype(F)
    d_rast[b, py, px, 0] = (dy * (a0 - a2)).sum(-1)
    d_rast[b, py, px, 1] = (dy * (a1 - a2)).sum(-1)
    d_attr = None
    if need_attr_grad:
        d_attr = np.zeros(attr.shape, dtype=F)
        b0 = rast[b, py, px, 0][:, None]
        b1 = rast[b, py, px, 1][:, None]
        b2 = (F(1.0) - b0) - b1
        if attr.ndim == 3:
            np.add.at(d_attr, (b, tri[t, 0]), dy * b0)
            np.add.at(d_attr, (b, tri[t, 1]), dy * b1)
            np.add.at(d_attr, (b, tri[t, 2]), dy * b2)
        else:
            np.add.at(d_attr, tri[t, 0], dy * b0)
            np.add.at(d_attr, tri[t, 1], dy * b1)
            np.add.at(d_attr, tri[t, 2], dy * b2)
    return d_rast, d_attr


# ----------------------------------------------------------------------------
# texture (filter_mode="linear", boundary_mode="wrap")


def _tex_taps(uv, Ht, Wt):
    u = uv[..., 0].astype(F)
    v = uv[..., 1].astype(F)
    u = u - np.floor(u)
    v = v - np.floor(v)
    u = u * F(Wt) - F(0.5)
    v = v * F(Ht) - F(0.5)
    iu0 = np.floor(u).astype(np.int64)
    iv0 = np.floor(v).astype(np.int64)
    fu = u - iu0.astype(F)
    fv = v - iv0.astype(F)
    iu1 = iu0 + 1
    iv1 = iv0 + 1
    iu0 = np.where(iu0 < 0, iu0 + Wt, iu0)
    iv0 = np.where(iv0 < 0, iv0 + Ht, iv0)
    iu1 = np.where(iu1 >= Wt, iu1 - Wt, iu1)
    iv1 = np.where(iv1 >= Ht, iv1 - Ht, iv1)
    return iu0, iv0, iu1, iv1, fu, fv


def texture_linear(tex, uv):
    """`dr.texture(tex, uv, filter_mode="linear")` with the default wrap boundary
    (`diffdope/diffdope.py:221-226`). tex [Ht,Wt,C] shared by
    the batch (the reference stacks B identical copies), uv [B,H,W,2]."""
    tex = np.asarray(tex, dtype=F)
    Ht, Wt, _ = tex.shape
    iu0, iv0, iu1, iv1, fu, fv = _tex_taps(uv, Ht, Wt)
    a00, a10 = tex[iv0, iu0], tex[iv0, iu1]
    a01, a11 = tex[iv1, iu0], tex[iv1, iu1]
    fu, fv = fu[..., None], fv[..., None]
    top = a00 + (a10 - a00) * fu
    bot = a01 + (a11 - a01) * fu
    return top + (bot - top) * fv


def texture_linear_grad_uv(tex, uv, d_out):
    """`TextureGradKernelLinear`: d_uv only (the texture itself has no grad in the
    reference: `Mesh.enable_gradients_texture` is dead code, `diffdope.py:1341,1361`)."""
    tex = np.asarray(tex, dtype=F)
    Ht, Wt, _ = tex.shape
    iu0, iv0, iu1, iv1, fu, fv = _tex_taps(uv, Ht, Wt)
    a00, a10 = tex[iv0, iu0], tex[iv0, iu1]
    a01, a11 = tex[iv1, iu0], tex[iv1, iu1]
    ad = (a11 + a00) - (a10 + a01)
    dy = d_out.astype(F)
    gu = (dy * ((a10 - a00) + fv[..., None] * ad)).sum(-1) * F(Wt)
    gv = (dy * ((a01 - a00) + fu[..., None] * ad)).sum(-1) * F(Ht)
    return np.stack([gu, gv], -1).astype(F)


def texture_linear_grad_tex(tex_shape, uv, d_out):
    """`TextureGradKernelLinear`, texture part: the bilinear weights of every lookup scattered to its four taps (what autograd
    sends into `tex` once `Mesh.enable_gradients_texture()` made it a parameter, `diffdope/diffdope.py:909-920`). One texture
    shared by the batch: the gradient is summed over it. Accumulated in float64, returned as float32."""
    Ht, Wt, C = tex_shape
    iu0, iv0, iu1, iv1, fu, fv = _tex_taps(uv, Ht, Wt)
    g = np.zeros((Ht * Wt, C), dtype=np.float64)
    dy = d_out.reshape(-1, C).astype(np.float64)
    fu, fv = fu.reshape(-1, 1).astype(np.float64), fv.reshape(-1, 1).astype(np.float64)
    for iv, iu, w in ((iv0, iu0, (1 - fu) * (1 - fv)), (iv0, iu1, fu * (1 - fv)), (iv1, iu0, (1 - fu) * fv), (iv1, iu1, fu * fv)):
        np.add.at(g, (iv * Wt + iu).reshape(-1), w * dy)
    return g.reshape(Ht, Wt, C).astype(F)


# ----------------------------------------------------------------------------
# texture, "linear-mipmap-linear" -- EXTENSION with no reference counterpart (the reference calls
# dr.texture with filter_mode="linear" only, `diffdope/diffdope.py:221-226`; BASELINE.json's
# north_star asks for a mipmapped sample). Definition:
#   * chain: level l+1 = 2x2 box filter of level l (a dimension of 1 stays 1), down to 1x1;
#   * level of detail per pixel: 0.5*log2 of the squared major axis (in level-0 texels) of the pixel
#     footprint in texture space, from the analytic screen derivatives of the interpolated uv
#     (the ellipse formula nvdiffrast's mip lookup uses), clamped to [0, levels-1];
#   * colour: linear blend of two bilinear wrap lookups at floor(lod) and floor(lod)+1;
#   * gradient: to uv through both lookups; the level of detail is a constant.


def build_mip_chain(tex, max_levels=0):
    """List of float32 levels [h,w,C], level 0 = tex."""
    lv = [np.asarray(tex, dtype=F)]
    while (lv[-1].shape[0] > 1 or lv[-1].shape[1] > 1) and (max_levels <= 0 or len(lv) < max_levels) and len(lv) < 15:
        t = lv[-1]
        h, w = t.shape[:2]
        dh, dw = max(h >> 1, 1), max(w >> 1, 1)
        y0 = np.minimum(2 * np.arange(dh), h - 1)
        y1 = np.minimum(2 * np.arange(dh) + 1, h - 1)
        x0 = np.minimum(2 * np.arange(dw), w - 1)
        x1 = np.minimum(2 * np.arange(dw) + 1, w - 1)
        a, b = t[y0][:, x0], t[y0][:, x1]
        c, d = t[y1][:, x0], t[y1][:, x1]
        lv.append((((a + b) + (c + d)) * F(0.25)).astype(F))
    return lv


def texture_lod(clip, tri, uv_attr, rast, tex_hw, n_levels):
    """Per-pixel level of detail [B,H,W] (0 at background). float64 inside: the value is
    continuous in its inputs, so no decision depends on its last bits."""
    B, H, W, _ = rast.shape
    Ht, Wt = tex_hw
    lod = np.zeros((B, H, W), dtype=F)
    tid = rast[..., 3].astype(np.int64) - 1
    b, py, px = np.nonzero(tid >= 0)
    if b.size == 0:
        return lod
    t = tid[b, py, px]
    c = np.asarray(clip, dtype=np.float64)
    c0, c1, c2 = c[b, tri[t, 0]], c[b, tri[t, 1]], c[b, tri[t, 2]]
    fx, fy = pixel_ndc(px, py, W, H)
    fx, fy = fx.astype(np.float64), fy.astype(np.float64)
    a0, a1, a2, (p0x, p0y, p1x, p1y, p2x, p2y) = shader_terms(c0, c1, c2, fx, fy)
    w0, w1, w2 = c0[:, 3], c1[:, 3], c2[:, 3]
    a0x, a0y = p1y * w2 - w1 * p2y, w1 * p2x - p1x * w2
    a1x, a1y = p2y * w0 - w2 * p0y, w2 * p0x - p2x * w0
    a2x, a2y = p0y * w1 - w0 * p1y, w0 * p1x - p0x * w1
    at = a0 + a1 + a2
    b0, b1 = a0 / at, a1 / at
    atx, aty = a0x + a1x + a2x, a0y + a1y + a2y
    xs, ys = 2.0 / W, 2.0 / H
    b0x, b0y = (a0x - b0 * atx) / at * xs, (a0y - b0 * aty) / at * ys
    b1x, b1y = (a1x - b1 * atx) / at * xs, (a1y - b1 * aty) / at * ys
    uv = np.asarray(uv_attr, dtype=np.float64)
    t0, t1, t2 = uv[tri[t, 0]], uv[tri[t, 1]], uv[tri[t, 2]]
    du0, du1 = t0[:, 0] - t2[:, 0], t1[:, 0] - t2[:, 0]
    dv0, dv1 = t0[:, 1] - t2[:, 1], t1[:, 1] - t2[:, 1]
    dudx, dvdx = (b0x * du0 + b1x * du1) * Wt, (b0x * dv0 + b1x * dv1) * Ht
    dudy, dvdy = (b0y * du0 + b1y * du1) * Wt, (b0y * dv0 + b1y * dv1) * Ht
    A, Bq, C = dudx * dudx + dvdx * dvdx, dudy * dudy + dvdy * dvdy, dudx * dudy + dvdx * dvdy
    major2 = 0.5 * (A + Bq) + np.sqrt(0.25 * (A - Bq) ** 2 + C * C)
    l = 0.5 * np.log2(np.maximum(major2, 1e-30))
    lod[b, py, px] = np.clip(l, 0.0, float(n_levels - 1)).astype(F)
    return lod


def _mip_levels_of(lod, n_levels):
    l0 = np.minimum(lod.astype(np.int64), n_levels - 1)
    f = (lod - l0.astype(F)).astype(F)
    use1 = (f > 0) & (l0 + 1 < n_levels)
    return l0, f, use1


def texture_mipmap(levels, uv, lod):
    """Trilinear lookup: levels from build_mip_chain, uv [B,H,W,2], lod [B,H,W]."""
    n = len(levels)
    l0, f, use1 = _mip_levels_of(lod, n)
    out = np.zeros(uv.shape[:-1] + (levels[0].shape[2],), dtype=F)
    for l in range(n):
        m0 = l0 == l
        if m0.any():
            out[m0] = texture_linear(levels[l], uv[m0])
        m1 = use1 & (l0 + 1 == l)
        if m1.any():
            c1 = texture_linear(levels[l], uv[m1])
            out[m1] = out[m1] + f[m1][:, None] * (c1 - out[m1])
    return out


def texture_mipmap_grad_uv(levels, uv, lod, d_out):
    n = len(levels)
    l0, f, use1 = _mip_levels_of(lod, n)
    g = np.zeros(uv.shape, dtype=F)
    for l in range(n):
        m0 = l0 == l
        if m0.any():
            g[m0] = texture_linear_grad_uv(levels[l], uv[m0], d_out[m0])
        m1 = use1 & (l0 + 1 == l)
        if m1.any():
            g1 = texture_linear_grad_uv(levels[l], uv[m1], d_out[m1])
            g[m1] = g[m1] + f[m1][:, None] * (g1 - g[m1])
    return g


# ----------------------------------------------------------------------------
# antialias


def _same_sign(a, b):
    return np.signbit(a) == np.signbit(b)


def _aa_pairs(tid, d):
    if d == 0:
        sel = tid[:, :, :-1] != tid[:, :, 1:]
    else:
        sel = tid[:, :-1, :] != tid[:, 1:, :]
    return np.nonzero(sel)


def _aa_analyse(rast, pos, tri, opp, b, py, px, d, H, W):
    """Vectorised `AntialiasFwdAnalysisKernel` for pixel pairs (px,py)-(px+1-d,py+d).

    Returns (valid, alpha, di, ds, tri_sel, qx, qy) per pair."""
    tid = rast[..., 3].astype(np.int64) - 1
    px1, py1 = px + (1 - d), py + d
    tri0, tri1 = tid[b, py, px], tid[b, py1, px1]
    z0, z1 = rast[b, py, px, 2], rast[b, py1, px1, 2]
    t = np.where(tri0 >= 0, tri0, tri1)
    both = (tri0 >= 0) & (tri1 >= 0)
    t = np.where(both, np.where(z0 < z1, tri0, tri1), t)
    is1 = t == tri1  # tri0 != tri1 for every work item
    qx = np.where(is1, px1, px)
    qy = np.where(is1, py1, py)
    vi = tri[t]  # [n,3]
    op = opp[t]
    P = pos[b[:, None], vi]  # [n,3,4]
    opi = np.where(op < 0, vi, op)
    O = pos[b[:, None], opi]
    xh, yh = F(W) * F(0.5), F(H) * F(0.5)
    one = F(1.0)
    fx = (qx.astype(F) + F(0.5)) - xh
    fy = (qy.astype(F) + F(0.5)) - yh
    with np.errstate(all="ignore"):
        w = one / P[:, :, 3]
        ow = one / O[:, :, 3]
        x = (P[:, :, 0] * w) * xh - fx[:, None]
        y = (P[:, :, 1] * w) * yh - fy[:, None]
        ox = (O[:, :, 0] * ow) * xh - fx[:, None]
        oy = (O[:, :, 1] * ow) * yh - fy[:, None]
        x0, x1, x2 = x[:, 0], x[:, 1], x[:, 2]
        y0, y1, y2 = y[:, 0], y[:, 1], y[:, 2]
        bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        a0 = (x1 - ox[:, 0]) * (y2 - oy[:, 0]) - (x2 - ox[:, 0]) * (y1 - oy[:, 0])
        a1 = (x2 - ox[:, 1]) * (y0 - oy[:, 1]) - (x0 - ox[:, 1]) * (y2 - oy[:, 1])
        a2 = (x0 - ox[:, 2]) * (y1 - oy[:, 2]) - (x1 - ox[:, 2]) * (y0 - oy[:, 2])
        s0, s1, s2 = _same_sign(a0, bb), _same_sign(a1, bb), _same_sign(a2, bb)
        anysil = s0 | s1 | s2
        if d:
            x0, y0 = y0, x0
            x1, y1 = y1, x1
            x2, y2 = y2, x2
        dx0, dx1, dx2 = x2 - x1, x0 - x2, x1 - x0
        dy0, dy1, dy2 = y2 - y1, y0 - y2, y1 - y0
        ds = np.where(is1, F(-1.0), F(1.0)).astype(F)
        d0 = ds * (x1 * dy0 - y1 * dx0)
        d1 = ds * (x2 * dy1 - y2 * dx1)
        d2 = ds * (x0 * dy2 - y0 * dx2)
        k0, k1, k2 = _same_sign(y1, y2), _same_sign(y2, y0), _same_sign(y0, y1)
        neg = F(-FLT_MAX)
        r0 = np.where(k0, neg, d0 / np.where(k0, one, dy0))
        r1 = np.where(k1, neg, d1 / np.where(k1, one, dy1))
        r2 = np.where(k2, neg, d2 / np.where(k2, one, dy2))
        g10, g20, g21 = r1 > r0, r2 > r0, r2 > r1
        di = np.where(g20 & g21, 2, np.where(g10, 1, 0))
        ady0 = np.where(k0, one, np.abs(dy0))
        ady1 = np.where(k1, one, np.abs(dy1))
        ady2 = np.where(k2, one, np.abs(dy2))
        ok0 = (di == 0) & s0 & (ady0 >= np.abs(dx0))
        ok1 = (di == 1) & s1 & (ady1 >= np.abs(dx1))
        ok2 = (di == 2) & s2 & (ady2 >= np.abs(dx2))
        dc = np.where(ok0, r0, np.where(ok1, r1, np.where(ok2, r2, neg)))
        eps = F(0.0625)
        valid = anysil & (dc > -eps) & (dc < one + eps)
        dcc = np.minimum(np.maximum(dc, F(0.0)), one)
        alpha = ds * (F(0.5) - dcc)
    return valid, alpha.astype(F), di, ds, t, qx, qy


def antialias(color, rast, pos, tri, opp):
    """`dr.antialias(color, rast, pos, tri)` (`diffdope/diffdope.py:214`).

    Returns (out, work) where `work` is the per-pair record the backward needs.
    Contributions are accumulated in a fixed order (horizontal pairs by (b,y,x), then
    vertical pairs) so a pixel sums left, right, lower, upper pair in that order."""
    color = np.asarray(color, dtype=F)
    pos = np.asarray(pos, dtype=F)
    B, H, W, C = color.shape
    out = color.copy()
    tid = rast[..., 3].astype(np.int64) - 1
    work = []
    flat = out.reshape(B * H * W, C)
    for d in (0, 1):
        b, py, px = _aa_pairs(tid, d)
        if b.size == 0:
            continue
        valid, alpha, di, ds, t, qx, qy = _aa_analyse(rast, pos, tri, opp, b, py, px, d, H, W)
        b, py, px = b[valid], py[valid], px[valid]
        alpha, di, ds, t, qx, qy = alpha[valid], di[valid], ds[valid], t[valid], qx[valid], qy[valid]
        pix0 = (b * H + py) * W + px
        pix1 = pix0 + (W if d else 1)
        c0 = color.reshape(B * H * W, C)[pix0]
        c1 = color.reshape(B * H * W, C)[pix1]
        tgt = np.where(alpha > 0, pix0, pix1)
        np.add.at(flat, tgt, alpha[:, None] * (c1 - c0))
        work.append(dict(d=d, b=b, pix0=pix0, pix1=pix1, tgt=tgt, alpha=alpha, di=di, ds=ds, t=t, qx=qx, qy=qy))
    return out, work


def antialias_grad(color, rast, pos, tri, work, d_out):
    """`AntialiasGradKernel`: d_color and d_pos (clip x,y,w of the active edge's two
    vertices). 1/(dy + copysign(1e-3, dy)) and the |alpha| >= 0.5 kill follow the
    published kernel."""
    color = np.asarray(color, dtype=F)
    pos = np.asarray(pos, dtype=F)
    B, H, W, C = color.shape
    V = pos.shape[1]
    g_color = d_out.astype(F).copy()
    g_pos = np.zeros((B, V, 4), dtype=F)
    gcf = g_color.reshape(B * H * W, C)
    dyf = d_out.astype(F).reshape(B * H * W, C)
    cf = color.reshape(B * H * W, C)
    for wk in work:
        d, b, pix0, pix1, tgt, alpha = wk["d"], wk["b"], wk["pix0"], wk["pix1"], wk["tgt"], wk["alpha"]
        dy = dyf[tgt]
        dd = (dy * (cf[pix1] - cf[pix0])).sum(-1)
        v0 = dy * alpha[:, None]
        np.add.at(gcf, pix0, -v0)
        np.add.at(gcf, pix1, v0)
        di, t, qx, qy = wk["di"], wk["t"], wk["qx"], wk["qy"]
        i1 = np.where(di < 2, di + 1, 0)
        i2 = np.where(i1 < 2, i1 + 1, 0)
        vi1 = tri[t, i1]
        vi2 = tri[t, i2]
        p1 = pos[b, vi1].copy()
        p2 = pos[b, vi2].copy()
        pxh, pyh = F(W) * F(0.5), F(H) * F(0.5)
        fx = (qx.astype(F) + F(0.5)) - pxh
        fy = (qy.astype(F) + F(0.5)) - pyh
        if d:
            p1[:, [0, 1]] = p1[:, [1, 0]]
            p2[:, [0, 1]] = p2[:, [1, 0]]
            pxh, pyh = pyh, pxh
            fx, fy = fy, fx
        with np.errstate(all="ignore"):
            w1 = F(1.0) / p1[:, 3]
            w2 = F(1.0) / p2[:, 3]
            x1 = p1[:, 0] * w1 * pxh - fx
            y1 = p1[:, 1] * w1 * pyh - fy
            x2 = p2[:, 0] * w2 * pxh - fx
            y2 = p2[:, 1] * w2 * pyh - fy
            dx = x2 - x1
            dyy = y2 - y1
            db = x1 * dyy - y1 * dx
            ep = np.copysign(F(1e-3), dyy)
            iy = F(1.0) / (dyy + ep)
            dby = db * iy
            iw1 = -w1 * iy * dd
            iw2 = w2 * iy * dd
            gp1x = iw1 * pxh * y2
            gp2x = iw2 * pxh * y1
            gp1y = iw1 * pyh * (dby - x2)
            gp2y = iw2 * pyh * (dby - x1)
            gp1w = -(p1[:, 0] * gp1x + p1[:, 1] * gp1y) * w1
            gp2w = -(p2[:, 0] * gp2x + p2[:, 1] * gp2y) * w2
        if d:
            gp1x, gp1y = gp1y, gp1x
            gp2x, gp2y = gp2y, gp2x
        kill = (np.abs(alpha) >= F(0.5)) | (dd == 0)
        z = np.zeros_like(gp1x)
        g1 = np.where(kill[:, None], F(0.0), np.stack([gp1x, gp1y, z, gp1w], -1)).astype(F)
        g2 = np.where(kill[:, None], F(0.0), np.stack([gp2x, gp2y, z, gp2w], -1)).astype(F)
        np.add.at(g_pos, (b, vi1), g1)
        np.add.at(g_pos, (b, vi2), g2)
    return g_color, g_pos
