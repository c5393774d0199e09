"""CPU: host-side logic, the C-ABI library's exported surface, API import without a GPU."""
import ctypes
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

import scene_util as su

ROOT = su.ROOT
LIB = os.path.join(ROOT, "diff-dope_b200", "diffdope", "_lib", "libddope_b200.so")
HEADER = os.path.join(ROOT, "include", "ddope_b200.h")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "diff-dope_b200", "csrc")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    assert os.path.exists(LIB)


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddope_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    names = _declared_symbols()
    assert len(names) >= 22
    for n in names:
        assert hasattr(lib, n), "libddope_b200.so does not export %s" % n
    lib.ddope_abi_version.restype = ctypes.c_int
    assert lib.ddope_abi_version() == 2


def test_library_is_sm100a_with_lineinfo():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_error_reporting_without_gpu_compute():
    lib = ctypes.CDLL(LIB)
    lib.ddope_last_error.restype = ctypes.c_char_p
    lib.ddope_xfm_fwd.restype = ctypes.c_int
    rc = lib.ddope_xfm_fwd(None, 1, 4, None, 1, 1, None, None)
    assert rc != 0 and b"null pointer" in lib.ddope_last_error()
    lib.ddope_scene_set_window.restype = ctypes.c_int
    assert lib.ddope_scene_set_window(None, 0, 0, 1, 1) != 0


def test_python_binding_loads_and_has_no_cpu_fallback():
    import diffdope as dd
    from diffdope import _native

    assert _native.lib() is not None
    with pytest.raises(RuntimeError):
        dd.xfm_points(torch.zeros(1, 4, 3), torch.eye(4)[None])  # CPU tensors are refused, not emulated
    out = dd.xfm_points(torch.rand(2, 5, 3), torch.rand(2, 4, 4), use_python=True)  # the reference's torch validation path
    assert out.shape == (2, 5, 4)


def test_public_names_match_reference_api():
    import diffdope as dd

    for name in ("DiffDope", "Scene", "Object3D", "Camera", "Mesh", "Image", "xfm_points", "xfm_vectors", "render_texture_batch",
                 "matrix_batch_44_from_position_quat", "opencv_2_opengl", "interpolate", "l1_rgb_with_mask", "l1_depth_with_mask",
                 "l1_mask", "dist_batch_lr", "find_crop", "make_grid", "make_grid_image", "make_grid_overlay_batch"):
        assert hasattr(dd, name), name
    for meth in ("run_optimization", "get_argmin", "get_pose", "render_img", "plot_losses", "make_animation", "set_batchsize", "add_loss_value", "cuda"):
        assert hasattr(dd.DiffDope, meth), meth


# ---- PLY ---------------------------------------------------------------------------------------


def test_ply_ascii_and_binary_roundtrip(tmp_path):
    from diffdope._ply import load_ply

    verts = np.array([[0, 0, 0, 0.1, 0.2, 255, 0, 0], [1, 0, 0, 0.3, 0.4, 0, 255, 0], [0, 1, 0, 0.5, 0.6, 0, 0, 255], [1, 1, 0, 0.7, 0.8, 9, 9, 9]], dtype=np.float64)
    header = ("ply\nformat %s 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nproperty float s\nproperty float t\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 2\nproperty list uchar int vertex_indices\nend_header\n")
    pa = tmp_path / "a.ply"
    with open(pa, "w") as f:
        f.write(header % "ascii")
        for v in verts:
            f.write("%g %g %g %g %g %d %d %d\n" % tuple(v))
        f.write("3 0 1 2\n4 0 1 3 2\n")
    pb = tmp_path / "b.ply"
    with open(pb, "wb") as f:
        f.write((header % "binary_little_endian").encode())
        for v in verts:
            f.write(struct.pack("<5f3B", *v[:5], *[int(x) for x in v[5:]]))
        f.write(struct.pack("<B3i", 3, 0, 1, 2))
        f.write(struct.pack("<B4i", 4, 0, 1, 3, 2))
    for p in (pa, pb):
        m = load_ply(str(p))
        assert np.allclose(m.vertices, verts[:, :3])
        assert np.allclose(m.uv, verts[:, 3:5])
        assert m.vertex_colors.tolist() == verts[:, 5:].astype(int).tolist()
        assert m.faces.tolist() == [[0, 1, 2], [0, 1, 3], [0, 3, 2]]  # quad fan-triangulated


def test_ply_binary_variants(tmp_path):
    """binary_big_endian, all-triangle faces (vectorised read) with an extra scalar and a per-face texcoord list, empty face element."""
    from diffdope._ply import load_ply

    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], dtype=np.float32)
    faces = [[0, 1, 2], [1, 3, 2], [0, 3, 1]]
    for bo, fmt in (("<", "binary_little_endian"), (">", "binary_big_endian")):
        header = ("ply\nformat %s 1.0\ncomment TextureFile missing.png\nelement vertex 4\nproperty double x\nproperty double y\nproperty double z\n"
                  "property float nx\nproperty float ny\nproperty float nz\nelement face 3\nproperty uchar flags\n"
                  "property list uchar uint vertex_index\nproperty list uchar float texcoord\nend_header\n" % fmt)
        p = tmp_path / (fmt + ".ply")
        with open(p, "wb") as f:
            f.write(header.encode())
            for v in verts:
                f.write(struct.pack(bo + "3d3f", *[float(x) for x in v], 0.0, 0.0, 1.0))
            for k, fc in enumerate(faces):
                f.write(struct.pack(bo + "BB3IB6f", k, 3, *fc, 6, *[0.1 * k + 0.01 * j for j in range(6)]))
        m = load_ply(str(p))
        assert np.array_equal(m.vertices, verts.astype(np.float64))
        assert np.array_equal(m.vertex_normals, np.tile([0.0, 0.0, 1.0], (4, 1)))
        assert m.faces.tolist() == faces and m.faces.dtype == np.int64
        assert m.texture_file == "missing.png" and m.texture_image is None and m.uv is None
    # no faces at all
    p = tmp_path / "points.ply"
    with open(p, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\n"
                b"element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        f.write(struct.pack("<3f", 1, 2, 3))
    m = load_ply(str(p))
    assert m.vertices.tolist() == [[1.0, 2.0, 3.0]] and m.faces.shape == (0, 3)
    with pytest.raises(ValueError):
        q = tmp_path / "bad.ply"
        q.write_bytes(b"plx\n")
        load_ply(str(q))


def test_quaternion_helpers_match_scipy():
    from scipy.spatial.transform import Rotation as R

    from diffdope._quat import quat_from_matrix, quat_mul, quat_to_matrix33, rotation_to_quat

    rng = np.random.default_rng(0)
    for _ in range(20):
        r = R.random(random_state=rng.integers(1 << 30))
        q = quat_from_matrix(r.as_matrix())
        s = r.as_quat()
        assert np.allclose(q, s, atol=1e-9) or np.allclose(q, -s, atol=1e-9)
        assert np.allclose(quat_to_matrix33(q * 3.0), r.as_matrix(), atol=1e-9)
        assert np.allclose(rotation_to_quat(list(r.as_matrix().reshape(-1))), q)
        r2 = R.random(random_state=rng.integers(1 << 30))
        assert np.allclose(quat_to_matrix33(quat_mul(q, quat_from_matrix(r2.as_matrix()))), (r * r2).as_matrix(), atol=1e-9)


def test_matrix_from_quat_matches_rotation():
    from scipy.spatial.transform import Rotation as R

    import diffdope as dd

    r = R.random(random_state=5)
    q = torch.tensor(r.as_quat()[None], dtype=torch.float32)
    p = torch.tensor([[1.0, 2.0, 3.0]])
    M = dd.matrix_batch_44_from_position_quat(q, p)[0].numpy()
    assert np.allclose(M[:3, :3], r.as_matrix(), atol=1e-6) and np.allclose(M[:3, 3], [1, 2, 3]) and np.allclose(M[3], [0, 0, 0, 1])


def test_compat_shims_and_config():
    import diffdope  # noqa: F401  (activates the shims if the real packages are missing)
    import hydra
    from omegaconf import OmegaConf

    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
    assert cfg.camera.fx == 1390.53 and cfg.hyperparameters.nb_iterations == 60 and cfg.losses.l1_mask is True
    assert set(cfg.keys()) == {"camera", "scene", "object3d", "losses", "hyperparameters", "render_images", "render"}  # "render" holds an extension key
    assert dict(**cfg.camera)["im_height"] == 1080
    assert len(cfg.object3d.rotation) == 9 and cfg.hyperparameters.learning_rates_bound[1] == 100
    assert hasattr(hydra, "main") and hasattr(hydra.core.hydra_config, "HydraConfig")


def test_hydra_shim_runs_a_main_with_overrides(tmp_path):
    script = tmp_path / "m.py"
    script.write_text(
        "import sys\nsys.path.insert(0, %r)\nimport diffdope\nimport hydra\nfrom omegaconf import DictConfig\n"
        "@hydra.main(version_base=None, config_path=%r, config_name='diffdope')\n"
        "def main(cfg: DictConfig):\n"
        "    import hydra as h\n"
        "    print('B', cfg.hyperparameters.batchsize, cfg.losses.l1_rgb_with_mask, h.core.hydra_config.HydraConfig.get().runtime.output_dir)\n"
        "main()\n" % (os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "configs"))
    )
    out = subprocess.run([sys.executable, str(script), "hyperparameters.batchsize=3", "losses.l1_rgb_with_mask=true"], capture_output=True, text=True, cwd=tmp_path)
    assert out.returncode == 0, out.stderr
    assert "B 3 True" in out.stdout and "outputs" in out.stdout


def test_reference_example_reaches_the_device_under_the_shims(tmp_path):
    """The reference's own examples/simple_scene.py, unmodified, with the in-repo package and the hydra / omegaconf /
    icecream stand-ins on the path: imports, config loading and `hydra.main` work, and without a GPU the first failure
    is CUDA initialisation inside `DiffDope.__post_init__` -- not an import or config error. (With a GPU the same
    script runs to the end: tests/test_gpu_api.py.)"""
    ref = "/root/reference/examples/simple_scene.py"
    if not os.path.exists(ref) or torch.cuda.is_available():
        pytest.skip("needs the reference tree and no GPU")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(su.ROOT, "diff-dope_b200"), os.path.join(su.ROOT, "diff-dope_b200", "compat")]))
    out = subprocess.run([sys.executable, ref, "hyperparameters.nb_iterations=2", "hydra.run.dir=%s" % tmp_path], capture_output=True, text=True,
                         cwd=su.ROOT, env=env, timeout=300)
    assert out.returncode != 0
    assert "NVIDIA" in out.stderr or "CUDA" in out.stderr, out.stderr[-1500:]
    assert "ImportError" not in out.stderr and "ModuleNotFoundError" not in out.stderr and "ConfigAttributeError" not in out.stderr
    assert "__post_init__" in out.stderr


def test_find_crop_centred_extension():
    """The better crop finder (extension; readme.md:30): centred, whole tiles, inside the image, contains the grown bounding box."""
    import diffdope as dd

    seg = torch.from_numpy(su.example_targets(0.5)["segmentation"])
    y0, x0, h, w = dd.find_crop_centred(seg)
    rows, cols = torch.nonzero(seg[..., 0] > 0).T
    top, bottom, left, right = int(rows.min()), int(rows.max()), int(cols.min()), int(cols.max())
    assert h == w and h % 32 == 0 and 0 <= y0 and y0 + h <= 540 and 0 <= x0 and x0 + w <= 960
    assert y0 <= top and x0 <= left and y0 + h > bottom and x0 + w > right
    assert abs((x0 + w / 2) - (left + right + 1) / 2) <= 1, "centred where the image allows it"
    old = dd.find_crop(seg)  # the reference's: top-left anchored, side = max extent of the grown box
    assert old[2] < h + 32
    # elongated region near a corner: shifted inside, not cut; empty mask: the whole image
    m = torch.zeros(100, 200)
    m[2:10, 150:199] = 1
    y0, x0, h, w = dd.find_crop_centred(m, percentage=0.1, multiple=32, min_size=32)
    assert (h, w) == (64, 64) and y0 == 0 and x0 + w <= 200 and x0 <= 150 and x0 + w >= 199
    assert dd.find_crop_centred(torch.zeros(50, 60)) == (0, 0, 50, 60)
    assert dd.find_crop_centred(torch.ones(40, 300), multiple=32) == (0, 0, 40, 300), "larger than the image: the image"


def test_shipped_reference_inputs_are_verbatim():
    """tests/golden/reference_examples/simple_scene.py and tests/golden/configs/diffdope.yaml are the reference's files, byte for byte."""
    pairs = [("tests/golden/reference_examples/simple_scene.py", "/root/reference/examples/simple_scene.py"),
             ("tests/golden/reference_examples/run_bop_scene.py", "/root/reference/examples/run_bop_scene.py"),
             ("tests/golden/hope_val_000001_scene_error_deg_040_trans_016.json", "/root/reference/data/hope/val/000001/scene_error_deg_040_trans_016.json"),
             ("tests/golden/configs/diffdope.yaml", "/root/reference/configs/diffdope.yaml")]
    if not os.path.exists(pairs[0][1]):
        pytest.skip("reference tree not on this box")
    for mine, ref in pairs:
        assert open(os.path.join(su.ROOT, mine), "rb").read() == open(ref, "rb").read(), mine


_VIZ_CHECK = r"""
import sys
sys.path.insert(0, sys.argv[1] + '/tests/golden'); sys.path.insert(0, sys.argv[1] + '/diff-dope_b200')
import numpy as np, torch
import diffdope as dd  # before the stand-ins take over the name 'diffdope'
import make_reference_vectors as mrv
ref = mrv.import_reference()
g = torch.Generator().manual_seed(0)
t = torch.rand(5, 3, 6, 7, generator=g) * 3 - 1
ok = []
for kw in (dict(), dict(normalize=True), dict(normalize=True, value_range=(0.0, 1.5)), dict(normalize=True, scale_each=True),
           dict(nrow=2, padding=1, pad_value=0.5), dict(normalize=True, scale_each=True, value_range=(-0.5, 2.0), nrow=3)):
    ok.append(torch.equal(ref.make_grid(t.clone(), **kw), dd.make_grid(t.clone(), **kw)))
t1, t3, t2 = torch.rand(4, 1, 5, 5, generator=g), torch.rand(3, 5, 5, generator=g), torch.rand(5, 5, generator=g)
ok.append(torch.equal(ref.make_grid(t1, nrow=2), dd.make_grid(t1, nrow=2)))
ok.append(torch.equal(ref.make_grid(t3), dd.make_grid(t3)))
ok.append(torch.equal(ref.make_grid(t2), dd.make_grid(t2)))
ok.append(torch.equal(ref.make_grid([t3, t3 * 0.5], nrow=1), dd.make_grid([t3, t3 * 0.5], nrow=1)))
ok.append(torch.equal(ref.make_grid(t, 2, 1, True, None, False, 0.25), dd.make_grid(t, 2, 1, True, None, False, 0.25)))  # positional order
imgs = [torch.rand(2, 4, 5, 3, generator=g) for _ in range(3)]
for w, h in ((1, 1), (2, 1), (2, 2), (3, 2)):
    ok.append(np.array_equal(ref.getimg_stack([i.clone() for i in imgs], w=w, h=h), dd.getimg_stack([i.clone() for i in imgs], w=w, h=h)))
dep = [torch.rand(2, 4, 5, generator=g) * 4 - 1 for _ in range(2)]
ok.append(np.array_equal(ref.getimg_stack([d.clone() for d in dep], depth=True, depth_max=3, w=2, h=1),
                         dd.getimg_stack([d.clone() for d in dep], depth=True, depth_max=3, w=2, h=1)))
print('VIZ', len(ok), all(ok), ok)
"""


def test_make_grid_options_and_getimg_stack_equal_the_reference_functions():
    """`make_grid` (normalize / value_range / scale_each, same positional order) and the legacy `getimg_stack` give
    bit for bit what the reference's own functions give (`diffdope/diffdope.py:277-309,337-460`), the reference module
    being imported in a subprocess with inert stand-ins for its missing third-party imports."""
    if not os.path.exists("/root/reference/diffdope/diffdope.py"):
        pytest.skip("reference tree not on this box")
    out = subprocess.run([sys.executable, "-c", _VIZ_CHECK, su.ROOT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    line = [l for l in out.stdout.splitlines() if l.startswith("VIZ")][-1]
    assert line.split()[1] == "16" and line.split()[2] == "True", line


def test_camera_and_image_loading_follow_reference():
    import diffdope as dd

    cam = dd.Camera(**su.CAMERA)
    assert np.allclose(cam.cam_proj.numpy(), su.projection_native())
    cam.set_batchsize(5)
    assert tuple(cam.cam_proj.shape) == (5, 4, 4)
    im = dd.Image(os.path.join(su.DATA, "scene", "seg.png"), img_resize=0.5)
    assert np.array_equal(im.img_tensor.numpy(), su.example_targets(0.5)["segmentation"])
    im.set_batchsize(4)
    assert tuple(im.img_tensor.shape) == (4, 540, 960, 3) and im.img_tensor.stride(0) == 0  # a view, not 4 copies
    d = dd.Image(os.path.join(su.DATA, "scene", "depth.png"), img_resize=0.5, depth=True)
    assert np.array_equal(d.img_tensor.numpy(), su.example_targets(0.5)["depth"])


def test_device_image_arithmetic_equals_the_host_pipeline():
    """`ddope_image_from_raw` (csrc/image.cu) is float64 arithmetic rounded once to float32: sample / divisor, and at
    img_resize = 0.5 the 2x2 area mean (((a+b)+c)+d) * 0.25 for colour, pixel (2y, 2x) for depth. This restates that
    arithmetic in numpy and checks it against the reference's cv2 pipeline (`Image.__post_init__`, diffdope.py:1122-1152)
    on the example images and on random even-sized images (SURVEY.md Appendix D pins 6-7); the kernel itself is compared
    with the host pipeline on the GPU (tests/test_gpu_wire.py)."""
    import cv2

    def kernel_arithmetic(raw, depth, divisor, half):
        f = (raw[::-1].astype(np.float64) if depth else raw[::-1, :, 2::-1].astype(np.float64)) / divisor
        if half:
            f = f[0::2, 0::2] if depth else (((f[0::2, 0::2] + f[0::2, 1::2]) + f[1::2, 0::2]) + f[1::2, 1::2]) * 0.25
        return f.astype(np.float32)

    sc = os.path.join(su.DATA, "scene")
    for name, depth in (("rgb.png", False), ("seg.png", False), ("depth.png", True)):
        raw = cv2.imread(os.path.join(sc, name), cv2.IMREAD_UNCHANGED if depth else cv2.IMREAD_COLOR)
        for resize in (1.0, 0.5):
            host = su.load_image(os.path.join(sc, name), resize, depth=depth)
            assert np.array_equal(kernel_arithmetic(raw, depth, 100.0 if depth else 255.0, resize == 0.5), host), (name, resize)
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 256, (54, 72, 3), dtype=np.uint8)
    f = cv2.flip(cv2.cvtColor(raw, cv2.COLOR_BGR2RGB) / 255.0, 0)
    assert np.array_equal(kernel_arithmetic(raw, False, 255.0, True), cv2.resize(f, (36, 27)).astype(np.float32))
    rawd = rng.integers(0, 65536, (54, 72), dtype=np.uint16)
    fd = cv2.flip(rawd / 100, 0)
    assert np.array_equal(kernel_arithmetic(rawd, True, 100.0, True), cv2.resize(fd, (36, 27), interpolation=cv2.INTER_NEAREST).astype(np.float32))


def test_mesh_and_object3d_follow_reference():
    import diffdope as dd

    m = dd.Mesh(os.path.join(su.DATA, "mesh", "AlphabetSoup.ply"), scale=0.01)
    a = su.example_mesh_arrays()
    assert m.has_textured_map and np.allclose(m.pos.numpy(), a["pos"], atol=1e-7) and np.array_equal(m.pos_idx.numpy(), a["tri"])
    assert np.allclose(m.uv.numpy(), a["uv"]) and tuple(m.tex.shape) == (2048, 2048, 3)
    m.set_batchsize(6)
    out = m()
    assert tuple(out["pos"].shape) == (6, 8240, 3) and tuple(out["tex"].shape) == (6, 2048, 2048, 3)
    assert out["tex"].stride(0) == 0, "texture must not be stacked B times"
    o = dd.Object3D(su.POSITION, su.ROTATION, batchsize=4, scale=0.01)
    q, t = su.example_pose()
    assert tuple(o.qx.shape) == (4,) and np.allclose([o.qx[0].item(), o.qy[0].item(), o.qz[0].item(), o.qw[0].item()], q, atol=1e-6)
    assert np.allclose([o.x[0].item(), o.y[0].item(), o.z[0].item()], t, atol=1e-6)
    assert len(list(o.parameters())) == 7


def test_obj_reader_matches_the_ply_reader_and_handles_seams(tmp_path):
    """Wavefront OBJ (what trimesh.load serves the reference for .obj models): the example mesh rewritten as OBJ + MTL loads to the
    arrays the PLY gives; a hand-written file covers separate v / vt indices (uv seam -> duplicated vertex), a quad, negative
    indices, `v//vn` corners and an untextured file with vertex colours; other extensions fail with a clear message."""
    import cv2

    import diffdope as dd
    from diffdope._obj import load_obj
    from diffdope._ply import load_ply

    src = os.path.join(su.DATA, "mesh", "AlphabetSoup.ply")
    ply = load_ply(src)
    obj = tmp_path / "soup.obj"
    with open(obj, "w") as f:
        f.write("mtllib soup.mtl\nusemtl m0\n")
        for v in ply.vertices:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for t in ply.uv:
            f.write("vt %r %r\n" % tuple(float(x) for x in t))
        for a, b, c in ply.faces:
            f.write("f %d/%d %d/%d %d/%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
    (tmp_path / "soup.mtl").write_text("newmtl m0\nKd 1 1 1\nmap_Kd -clamp on %s\n" % os.path.join(su.DATA, "mesh", "AlphabetSoup.png"))
    o = load_obj(str(obj))
    # every vertex of the example is used by a face, but not in index order: compare through the faces
    assert o.faces.shape == ply.faces.shape
    assert np.array_equal(o.vertices[o.faces], ply.vertices[ply.faces]) and np.array_equal(o.uv[o.faces], ply.uv[ply.faces])
    assert np.array_equal(o.texture_image, ply.texture_image)
    m_obj, m_ply = dd.Mesh(str(obj), scale=0.01), dd.Mesh(src, scale=0.01)
    assert m_obj.has_textured_map and torch.equal(m_obj.tex, m_ply.tex)
    assert torch.equal(m_obj.pos[m_obj.pos_idx.long()], m_ply.pos[m_ply.pos_idx.long()]) and torch.equal(m_obj.uv[m_obj.uv_idx.long()], m_ply.uv[m_ply.pos_idx.long()])

    tex = tmp_path / "t.png"
    cv2.imwrite(str(tex), np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3))
    (tmp_path / "q.mtl").write_text("newmtl skin\nmap_Kd t.png\n")
    (tmp_path / "q.obj").write_text(
        "# a quad and a triangle sharing an edge, the shared corners with different uv on the two faces\n"
        "mtllib q.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 0 0\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\nvt 0.25 0.75\nvn 0 0 1\n"
        "usemtl skin\nf 1/1/1 2/2/1 3/3/1 4/4/1\nf -4/5/1 -1/6/1 3/3/1\n")
    q = load_obj(str(tmp_path / "q.obj"))
    assert q.faces.tolist() == [[0, 1, 2], [0, 2, 3], [4, 5, 2]], "fan triangulation, corners re-indexed in order of first appearance"
    assert q.vertices.shape == (6, 3) and np.array_equal(q.vertices[4], [1, 0, 0]) and np.array_equal(q.vertices[5], [2, 0, 0])
    assert np.array_equal(q.uv[4], [0.5, 0.5]) and np.array_equal(q.uv[1], [1, 0]) and np.array_equal(q.vertex_normals[5], [0, 0, 1])
    assert q.texture_image.shape == (4, 4, 3) and q.texture_image[0, 0].tolist() == [2, 1, 0], "RGB, row 0 = top row of the file"

    (tmp_path / "c.obj").write_text("v 0 0 0 1 0 0\nv 1 0 0 0 1 0\nv 0 1 0 0 0 1\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    c = load_obj(str(tmp_path / "c.obj"))
    assert c.uv is None and c.faces.tolist() == [[0, 1, 2]] and c.vertex_colors.tolist() == [[255, 0, 0], [0, 255, 0], [0, 0, 255]]
    mc = dd.Mesh(str(tmp_path / "c.obj"), scale=1.0)
    assert not mc.has_textured_map and tuple(mc.vtx_color.shape) == (3, 3)
    with pytest.raises(ValueError, match="unsupported mesh format"):
        dd.Mesh(str(tmp_path / "model.glb"), scale=1.0)


def test_host_pose_matrix_is_bit_equal_to_the_differentiable_one():
    """The numpy float32 builder behind the lazily made 'mtx' of a result equals matrix_batch_44_from_position_quat (itself pinned
    bit-equal to the reference's function by the golden vectors) bit for bit, on unit and non-unit quaternions."""
    from diffdope.diffdope import _matrix_batch_44_np, matrix_batch_44_from_position_quat

    g = torch.Generator().manual_seed(7)
    for scale in (1.0, 3.7, 1e-3):
        q = torch.randn(257, 4, generator=g)
        q = q / torch.norm(q, dim=-1, keepdim=True) * scale
        p = torch.randn(257, 3, generator=g) * 50.0
        a = matrix_batch_44_from_position_quat(q, p).numpy()
        b = _matrix_batch_44_np(q.numpy(), p.numpy())
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


def test_per_hypothesis_start_poses_survive_refills():
    """Extension: set_pose with [B,3] / [B,4] arrays. The poses must survive set_batchsize / reset_pose (which refill the parameters),
    a batch size they do not fit must fail loudly, and the representative quaternion of (q, -q) is q, not NaN."""
    import diffdope as dd

    o = dd.Object3D(su.POSITION, su.ROTATION, batchsize=2, scale=0.01)
    q = np.array([0.1, 0.2, 0.3, 0.9]); q /= np.linalg.norm(q)
    qs = np.stack([q, -q, q])
    ps = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]])
    o.set_pose(ps, qs, opencv2opengl=False)
    def rows():
        qq, tt = o.pose_tensors()
        return qq.numpy(), tt.numpy()
    qq, tt = rows()
    assert qq.shape == (3, 4) and np.allclose(qq, qs, atol=1e-7) and np.allclose(tt, ps)
    assert np.all(np.isfinite(o._rotation)) and abs(abs(float(np.dot(o._rotation, q))) - 1.0) < 1e-6
    o.set_batchsize(3)
    qq, tt = rows()
    assert np.allclose(qq, qs, atol=1e-7) and np.allclose(tt, ps), "set_batchsize discarded the per-hypothesis poses"
    with torch.no_grad():
        o.x.add_(1.0)
    o.reset_pose()
    assert np.allclose(rows()[1], ps)
    with pytest.raises(ValueError):
        o.set_batchsize(5)
    o.set_pose([1.0, 2.0, 3.0], list(q), batchsize=5, opencv2opengl=False)  # one pose again: any batch size
    o.set_batchsize(7)
    assert rows()[0].shape == (7, 4)


def test_shard_bounds_cover_every_hypothesis_once():
    from diffdope import _dist

    for B in (1, 7, 8, 64, 65):
        for ws in (1, 2, 3, 8):
            seen = []
            for r in range(ws):
                lo, hi = _dist.shard_bounds(B, r, ws)
                seen += list(range(lo, hi))
            assert seen == list(range(B))


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    import json

    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "hyp*iter/s"


def test_mesh_orientation_host_function_matches_oracle():
    """ddope_mesh_orientation (pure host code in the C-ABI library) and oracle.nvdr.closed_mesh_orientation classify
    meshes identically: closed / inside-out / open / inconsistently wound / seam-duplicated."""
    from oracle import nvdr

    lib = ctypes.CDLL(LIB)
    lib.ddope_mesh_orientation.restype = ctypes.c_int
    lib.ddope_mesh_orientation.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]

    def c_side(pos, tri):
        p = np.ascontiguousarray(pos, dtype=np.float32)
        t = np.ascontiguousarray(tri, dtype=np.int32)
        return lib.ddope_mesh_orientation(p.ctypes.data, p.shape[0], t.ctypes.data, t.shape[0])

    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float32) * 0.5
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], dtype=np.int32)
    flipped = f.copy()
    flipped[3] = flipped[3, ::-1]
    v2 = np.concatenate([v, v[:1]])
    f2 = f.copy()
    f2[0, 0] = 8
    arr = su.example_mesh_arrays()
    cases = [(v, f, 1), (v, f[:, ::-1], -1), (v, f[:-1], 0), (v, flipped, 0), (v2, f2, 1), (arr["pos"], arr["tri"], 1), (arr["pos"], arr["tri"][:-1], 0)]
    for pos, tri, want in cases:
        assert c_side(pos, tri) == want == nvdr.closed_mesh_orientation(pos, tri)
    assert c_side(v, np.array([[0, 1, 99]], np.int32)) == 0  # out-of-range index: not classified, no crash


def test_baseline_config_workloads_are_well_formed():
    """tests/workloads.py: the stand-in workloads of BASELINE configs 3-5 have the sizes SURVEY.md 8(d) states, closed and
    consistently oriented meshes (so the back-face rule applies), in-range uv and poses in front of the camera."""
    import workloads as wl
    from oracle import nvdr

    c4 = wl.config4()
    assert c4["pos"].shape == (5002, 3) and c4["tri"].shape == (10000, 3) and (c4["H"], c4["W"]) == (540, 720) and c4["B"] == 256
    assert nvdr.closed_mesh_orientation(c4["pos"], c4["tri"]) == 1
    assert np.allclose(np.abs(c4["pos"]).max(0), [0.5, 0.5, 0.3], atol=1e-3)
    ang = 2 * np.degrees(np.arccos(abs(float(np.dot(c4["q0"], c4["q_gt"])))))
    assert abs(ang - 10.0) < 1e-3 and abs(np.linalg.norm(c4["t0"] - c4["t_gt"]) - 0.04) < 1e-6
    c5 = wl.config5(tex_size=64)
    assert c5["pos"].shape == (25002, 3) and c5["tri"].shape == (50000, 3) and (c5["H"], c5["W"]) == (1024, 1024) and c5["B"] == 1024
    assert nvdr.closed_mesh_orientation(c5["pos"], c5["tri"]) == 1
    assert c5["uv"].min() >= 0 and c5["uv"].max() <= 1 and c5["tex"].shape == (64, 64, 3) and 0.3 < c5["tex"].mean() < 0.7
    assert c5["losses"]["l1_edge"] and c5["t_gt"][2] < 0
    c3 = wl.config3()
    assert len(c3["objects"]) == 8 and (c3["H"], c3["W"]) == (480, 640) and c3["B"] == 128
    scales = [np.abs(o["pos"]).max() for o in c3["objects"]]
    assert np.allclose(np.array(scales) / scales[0], np.linspace(0.6, 1.3, 8) / 0.6, rtol=1e-5)
    for o in c3["objects"]:
        assert o["t_gt"][2] < -5 and abs(np.linalg.norm(o["q_gt"]) - 1) < 1e-5 and abs(np.linalg.norm(o["q0"]) - 1) < 1e-5
    # cycled copies are shifted so that no two of the eight objects coincide
    ts = np.array([o["t_gt"] for o in c3["objects"]])
    assert min(np.linalg.norm(ts[i] - ts[j]) for i in range(8) for j in range(i)) > 0.1
    tex = wl.procedural_texture(32, seed=0)
    assert np.array_equal(tex, wl.procedural_texture(32, seed=0)) and tex.dtype == np.float32


def test_bench_byte_model_matches_survey():
    """bench.py's roofline numerator is SURVEY.md 8(d)'s algorithmic byte model for config 2: 23.99 MB per
    hypothesis-iteration, split into the terms the pixel pass (17.10 MB) and the raster kernel (6.89 MB) move."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    m = bench.survey_bytes_per_hit()
    assert abs(m["total"] / 1e6 - 23.99) < 0.01
    assert abs(m["pixel_kernel"] / 1e6 - 17.10) < 0.01 and abs(m["raster_kernel"] / 1e6 - 6.89) < 0.01
    assert abs(m["pixel_kernel"] + m["raster_kernel"] - m["total"]) < 1
    assert bench.B_PER_GPU == 64 and bench.WINDOW == 640 and bench.METRIC.startswith("pose-hypotheses")
    sched = bench.lr_schedule(200)
    assert abs(sched[0] - 2.0) < 1e-12 and abs(sched[-1] - 0.2) < 1e-12
    # the other configurations' figures of SURVEY.md 8(d): 3.95 MB (config 1, default losses), 17.80 (3), 15.73 (4), 105.96 (5)
    assert abs(bench.survey_bytes_per_hit(P=320 * 320, c=0.24, rgb=False, depth=False)["total"] / 1e6 - 3.95) < 0.01
    assert abs(bench.survey_bytes_per_hit(P=640 * 480, c=0.05)["total"] / 1e6 - 17.80) < 0.03
    assert abs(bench.survey_bytes_per_hit(V=5002, T=10000, P=720 * 540, c=0.05, rgb=False, textured=False)["total"] / 1e6 - 15.73) < 0.01
    assert abs(bench.survey_bytes_per_hit(V=25002, T=50000, P=1024 * 1024, c=0.5)["total"] / 1e6 - 105.96) < 0.01
    assert len(bench.source_sha()) == 16 and bench.workload_config(4, "strong")["hypotheses_per_gpu"] == 64
