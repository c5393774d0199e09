"""Minimal Wavefront OBJ mesh reader.

Serves `Mesh` for `.obj` models the way `trimesh.load(path, force="mesh")` serves the reference
(`diffdope/diffdope.py:784-842`): positions `v`, texture coordinates `vt`, normals `vn`, faces `f` with the
`v`, `v/vt`, `v//vn` and `v/vt/vn` corner forms (negative = relative indices, polygons fan-triangulated), and the
diffuse texture `map_Kd` of the material the faces use (`mtllib` / `usemtl`).

OBJ indexes positions and texture coordinates separately (a vertex on a uv seam has one position and several
uv's). The renderer -- like nvdiffrast's `dr.interpolate(uv, rast, uv_idx)` with `uv_idx == pos_idx`, which is how
the reference's `Mesh` feeds it -- wants ONE index per corner, so corners are re-indexed by their distinct
(v, vt, vn) triple, in order of first appearance (trimesh does the same "unmerge"; the vertex order may differ
from trimesh's, the set of triangles, and therefore every rendered pixel, does not). A file whose faces carry no
`vt` keeps its positions exactly as stored.
"""
import os

import numpy as np

from ._ply import PlyMesh


def _mtl_textures(path):
    """material name -> absolute path of its map_Kd image (last token of the statement: options are skipped)."""
    out, cur = {}, None
    try:
        with open(path, "r", errors="replace") as f:
            for line in f:
                tok = line.split()
                if not tok or tok[0].startswith("#"):
                    continue
                if tok[0] == "newmtl":
                    cur = " ".join(tok[1:])
                elif tok[0].lower() == "map_kd" and cur is not None and len(tok) >= 2:
                    out[cur] = os.path.join(os.path.dirname(os.path.abspath(path)), tok[-1])
    except OSError:
        pass
    return out


def load_obj(path, load_texture=True):
    """-> PlyMesh-shaped container: vertices [V,3] f64, faces [T,3] i64, uv [V,2] | None, vertex_normals | None,
    vertex_colors (uint8, from the `v x y z r g b` extension) | None, texture_image uint8 [H,W,3] RGB | None."""
    pos, col, tex, nrm = [], [], [], []
    corners = []        # per face: list of (v, vt, vn) with -1 for absent, 0-based
    materials = []      # material in force when the face was read
    mtllibs, cur_mtl = [], None
    with open(path, "r", errors="replace") as f:
        for raw in f:
            tok = raw.split()
            if not tok or tok[0].startswith("#"):
                continue
            k = tok[0]
            if k == "v":
                pos.append([float(tok[1]), float(tok[2]), float(tok[3])])
                if len(tok) >= 7:
                    col.append([float(tok[4]), float(tok[5]), float(tok[6])])
            elif k == "vt":
                tex.append([float(tok[1]), float(tok[2]) if len(tok) > 2 else 0.0])
            elif k == "vn":
                nrm.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif k == "f":
                face = []
                for c in tok[1:]:
                    parts = c.split("/")
                    tri = []
                    for j, n in enumerate((len(pos), len(tex), len(nrm))):
                        s = parts[j] if j < len(parts) else ""
                        if s == "":
                            tri.append(-1)
                        else:
                            i = int(s)
                            tri.append(i - 1 if i > 0 else n + i)
                    face.append(tuple(tri))
                if len(face) >= 3:
                    corners.append(face)
                    materials.append(cur_mtl)
            elif k == "mtllib":
                mtllibs.append(" ".join(tok[1:]))
            elif k == "usemtl":
                cur_mtl = " ".join(tok[1:])
    if not pos:
        raise ValueError("%s: no vertices (not a Wavefront OBJ file?)" % path)
    pos = np.asarray(pos, dtype=np.float64)
    if any(c[0] < 0 or c[0] >= len(pos) for face in corners for c in face):
        raise ValueError("%s: face index out of range" % path)

    mesh = PlyMesh()
    have_vt = bool(tex) and all(c[1] >= 0 for face in corners for c in face) and bool(corners)
    have_vn = bool(nrm) and all(c[2] >= 0 for face in corners for c in face) and bool(corners)
    tris = []
    if have_vt:
        # one vertex per distinct (v, vt, vn) corner, first appearance first
        key_of, vi, ti, ni = {}, [], [], []
        for face in corners:
            idx = []
            for c in face:
                key = (c[0], c[1], c[2] if have_vn else -1)
                j = key_of.get(key)
                if j is None:
                    j = key_of[key] = len(vi)
                    vi.append(c[0]); ti.append(c[1]); ni.append(c[2])
                idx.append(j)
            for k in range(1, len(idx) - 1):
                tris.append((idx[0], idx[k], idx[k + 1]))
        vi = np.asarray(vi, dtype=np.int64)
        mesh.vertices = pos[vi]
        mesh.uv = np.asarray(tex, dtype=np.float64)[np.asarray(ti, dtype=np.int64)]
        if have_vn:
            mesh.vertex_normals = np.asarray(nrm, dtype=np.float64)[np.asarray(ni, dtype=np.int64)]
        if len(col) == len(pos):
            mesh.vertex_colors = np.clip(np.rint(np.asarray(col)[vi] * 255.0), 0, 255).astype(np.uint8)
    else:
        for face in corners:
            for k in range(1, len(face) - 1):
                tris.append((face[0][0], face[k][0], face[k + 1][0]))
        mesh.vertices = pos
        if have_vn and all(c[2] == c[0] for face in corners for c in face) and len(nrm) == len(pos):
            mesh.vertex_normals = np.asarray(nrm, dtype=np.float64)
        if len(col) == len(pos):
            mesh.vertex_colors = np.clip(np.rint(np.asarray(col) * 255.0), 0, 255).astype(np.uint8)
    mesh.faces = np.asarray(tris, dtype=np.int64).reshape(-1, 3)

    if have_vt and load_texture:
        base = os.path.dirname(os.path.abspath(path))
        table = {}
        for lib in mtllibs:
            table.update(_mtl_textures(os.path.join(base, lib)))
        used = [m for m in dict.fromkeys(materials) if m in table]
        if len(used) > 1 and len({table[m] for m in used}) > 1:
            raise ValueError("%s: faces use several textured materials (%s); one texture per mesh is supported, as in the reference's Mesh"
                             % (path, ", ".join(used)))
        cand = table[used[0]] if used else (next(iter(table.values())) if len(table) == 1 else None)
        if cand is not None and os.path.exists(cand):
            import cv2

            im = cv2.imread(cand, cv2.IMREAD_COLOR)
            if im is not None:
                mesh.texture_file = os.path.basename(cand)
                mesh.texture_image = np.ascontiguousarray(im[:, :, ::-1])
    return mesh
