"""Minimal PLY mesh reader (ASCII, binary_little_endian and binary_big_endian).

Replaces the `trimesh.load(path, force="mesh")` call of the reference
(`diffdope/diffdope.py:784-842`) for the asset format its examples use: a PLY
with per-vertex position, optional normals, optional per-vertex uv
(`texture_u/texture_v` or `s/t`), optional per-vertex colour and a texture
named by a `comment TextureFile <name>` header line.

Vertices are kept exactly as stored (no merging), faces with more than three
corners are fan-triangulated.
"""
import os

import numpy as np

_PLY_DTYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1",
    "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


class PlyMesh:
    """Plain container: vertices [V,3] f64, faces [T,3] i64, optional extras."""

    def __init__(self):
        self.vertices = None
        self.faces = None
        self.vertex_normals = None
        self.uv = None
        self.vertex_colors = None  # uint8 [V,3 or 4]
        self.texture_file = None
        self.texture_image = None  # uint8 [H,W,3] RGB, row 0 = top row of the file


def _parse_header(f):
    line = f.readline().strip()
    if line != b"ply":
        raise ValueError("not a PLY file")
    fmt = None
    elements = []  # [name, count, [(prop_name, kind, types...)]]
    texture_file = None
    while True:
        line = f.readline()
        if not line:
            raise ValueError("unexpected end of PLY header")
        tok = line.decode("ascii", "replace").strip().split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "comment":
            if len(tok) >= 3 and tok[1].lower() == "texturefile":
                texture_file = " ".join(tok[2:])
        elif tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1][2].append((tok[4], "list", tok[2], tok[3]))
            else:
                elements[-1][2].append((tok[2], "scalar", tok[1]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError("unsupported PLY format %r" % fmt)
    return fmt, elements, texture_file


def _triangulate(polys):
    tris = []
    for p in polys:
        for k in range(1, len(p) - 1):
            tris.append((p[0], p[k], p[k + 1]))
    return np.asarray(tris, dtype=np.int64).reshape(-1, 3)


def _read_binary_lists(f, count, props, bo):
    """Rows of a binary element that holds list properties. Meshes store a constant number of corners per face, so the
    whole element is first tried as one fixed-size record array (the count fields must all equal the first row's);
    mixed polygon sizes fall back to a row-by-row read."""
    lists = {p[0]: [] for p in props if p[1] == "list"}
    if count == 0:
        return lists
    start = f.tell()
    # sizes of the first row
    first, fields = [], []
    for i, p in enumerate(props):
        if p[1] == "list":
            cdt = np.dtype(bo + _PLY_DTYPES[p[2]])
            n = int(np.frombuffer(f.read(cdt.itemsize), dtype=cdt)[0])
            f.seek(np.dtype(_PLY_DTYPES[p[3]]).itemsize * n, 1)
            first.append(n)
            fields += [("c%d" % i, cdt), ("v%d" % i, bo + _PLY_DTYPES[p[3]], (n,))]
        else:
            f.seek(np.dtype(_PLY_DTYPES[p[2]]).itemsize, 1)
            fields.append(("s%d" % i, bo + _PLY_DTYPES[p[2]]))
    f.seek(start)
    dt = np.dtype(fields)
    buf = f.read(dt.itemsize * count)
    if len(buf) == dt.itemsize * count:
        arr = np.frombuffer(buf, dtype=dt, count=count)
        k = 0
        uniform = True
        for i, p in enumerate(props):
            if p[1] == "list":
                uniform = uniform and bool(np.all(arr["c%d" % i] == first[k]))
                k += 1
        if uniform:
            for i, p in enumerate(props):
                if p[1] == "list":
                    lists[p[0]] = arr["v%d" % i].astype(np.int64 if np.dtype(_PLY_DTYPES[p[3]]).kind in "iu" else np.float64).tolist()
            return lists
    f.seek(start)
    for _ in range(count):
        for p in props:
            if p[1] == "list":
                cdt = np.dtype(bo + _PLY_DTYPES[p[2]])
                idt = np.dtype(bo + _PLY_DTYPES[p[3]])
                n = int(np.frombuffer(f.read(cdt.itemsize), dtype=cdt)[0])
                vals = np.frombuffer(f.read(idt.itemsize * n), dtype=idt)
                lists[p[0]].append([int(v) for v in vals] if idt.kind in "iu" else [float(v) for v in vals])
            else:
                f.seek(np.dtype(_PLY_DTYPES[p[2]]).itemsize, 1)
    return lists


def load_ply(path, load_texture=True):
    mesh = PlyMesh()
    with open(path, "rb") as f:
        fmt, elements, texture_file = _parse_header(f)
        data = {}
        for name, count, props in elements:
            has_list = any(p[1] == "list" for p in props)
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(count)]
                if not has_list:
                    arr = np.asarray(rows, dtype=np.float64).reshape(count, len(props))
                    data[name] = {p[0]: arr[:, i] for i, p in enumerate(props)}
                else:
                    lists = {p[0]: [] for p in props if p[1] == "list"}
                    for r in rows:
                        c = 0
                        for p in props:
                            if p[1] == "list":
                                n = int(r[c])
                                lists[p[0]].append([int(float(v)) for v in r[c + 1:c + 1 + n]])
                                c += 1 + n
                            else:
                                c += 1
                    data[name] = lists
            else:
                bo = "<" if fmt == "binary_little_endian" else ">"
                if not has_list:
                    dt = np.dtype([(p[0], bo + _PLY_DTYPES[p[2]]) for p in props])
                    arr = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                    data[name] = {p[0]: arr[p[0]].astype(np.float64) for p in props}
                else:
                    data[name] = _read_binary_lists(f, count, props, bo)

    v = data["vertex"]
    mesh.vertices = np.stack([v["x"], v["y"], v["z"]], axis=1)
    if all(k in v for k in ("nx", "ny", "nz")):
        mesh.vertex_normals = np.stack([v["nx"], v["ny"], v["nz"]], axis=1)
    for ku, kv in (("texture_u", "texture_v"), ("s", "t"), ("u", "v")):
        if ku in v and kv in v:
            mesh.uv = np.stack([v[ku], v[kv]], axis=1)
            break
    if all(k in v for k in ("red", "green", "blue")):
        cols = [v["red"], v["green"], v["blue"]]
        if "alpha" in v:
            cols.append(v["alpha"])
        mesh.vertex_colors = np.stack(cols, axis=1).astype(np.uint8)

    face = data.get("face", {})
    idx_key = next((k for k in ("vertex_indices", "vertex_index") if k in face), None)
    if idx_key is None:
        mesh.faces = np.zeros((0, 3), dtype=np.int64)
    else:
        polys = face[idx_key]
        if all(len(p) == 3 for p in polys):
            mesh.faces = np.asarray(polys, dtype=np.int64).reshape(-1, 3)
        else:
            mesh.faces = _triangulate(polys)

    mesh.texture_file = texture_file
    if load_texture and texture_file is not None and mesh.uv is not None:
        tex_path = os.path.join(os.path.dirname(os.path.abspath(path)), texture_file)
        if os.path.exists(tex_path):
            import cv2

            im = cv2.imread(tex_path, cv2.IMREAD_COLOR)
            if im is not None:
                mesh.texture_image = np.ascontiguousarray(im[:, :, ::-1])
    return mesh
