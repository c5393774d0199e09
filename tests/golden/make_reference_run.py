"""Run the REFERENCE's own `DiffDope.run_optimization` loop, unmodified, on the CPU of the build container and store
what it produces as a golden fixture (the reference tree does not exist on the GPU box).

The reference module (/root/reference/diffdope/diffdope.py) needs packages that are absent here. They are replaced
as follows -- nothing in the reference file is edited:
  nvdiffrast.torch   the four ops (rasterize / interpolate / texture / antialias) are served by the oracle's
                     restatements (oracle/refpath.py autograd wrappers over oracle/nvdr.py), without back-face
                     culling (a GL context does not cull). So this fixture pins everything AROUND those ops: the render
                     graph's wiring and conventions, the losses, logging keys, schedule, optimiser, argmin, get_pose.
  diffdope (dd.*)    dd.xfm_points = the oracle's xfm (= the reference's own use_python formula, ops.py:137-141);
                     dd.l1_* = the reference module's own loss functions
  trimesh            `trimesh.load` returns the arrays of the in-repo PLY reader (the reference's Mesh.__init__ then
                     applies its own scaling, uv flip and /255)
  pyrr               only `pyrr.Quaternion(list)` is reached (pose given as an OpenGL quaternion, opencv2opengl=False)
  hydra, omegaconf, icecream, imageio, matplotlib    inert stand-ins / the repo's config stand-in
  .cuda()            no-op (CPU run)

    python tests/golden/make_reference_run.py   ->  tests/golden/reference_run.npz
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "diff-dope_b200", "compat"), os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import make_reference_vectors as mrv  # noqa: E402  (the stand-in machinery)
import scene_util as su  # noqa: E402
from oracle import nvdr, refpath  # noqa: E402

RESIZE, B, NB_ITER = 0.25, 2, 3


def install_shims():
    from diffdope._ply import load_ply  # the in-repo PLY reader (diffdope here is still the repo's package)

    # --- nvdiffrast.torch backed by the oracle ------------------------------------------------------------------
    dr = types.ModuleType("nvdiffrast.torch")
    opp_cache = {}

    def tri_np(t):
        return t.detach().cpu().numpy().astype(np.int64)

    def rasterize(glctx, pos, tri, resolution):
        r = refpath._Rasterize.apply(pos, tri_np(tri), int(resolution[0]), int(resolution[1]), None)
        return r, torch.zeros_like(r)

    def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
        return refpath._Interpolate.apply(attr, rast, tri_np(tri)), None

    def texture(tex, uv, uv_da=None, filter_mode="linear"):
        assert filter_mode == "linear"
        return refpath._TextureLinear.apply(tex[0] if tex.dim() == 4 else tex, uv)

    def antialias(color, rast, pos, tri):
        t = tri_np(tri)
        key = t.tobytes()
        if key not in opp_cache:
            opp_cache[key] = nvdr.build_edge_opposites(t)
        return refpath._Antialias.apply(color, rast, pos, t, opp_cache[key])

    dr.RasterizeGLContext = lambda *a, **k: object()
    dr.rasterize, dr.interpolate, dr.texture, dr.antialias = rasterize, interpolate, texture, antialias
    nv = types.ModuleType("nvdiffrast")
    nv.torch = dr
    sys.modules["nvdiffrast"], sys.modules["nvdiffrast.torch"] = nv, dr

    # --- trimesh: arrays from the in-repo PLY reader ------------------------------------------------------------
    tm = types.ModuleType("trimesh")
    tm.visual = types.ModuleType("trimesh.visual")
    tm.visual.texture = types.ModuleType("trimesh.visual.texture")

    class TextureVisuals:
        pass

    tm.visual.texture.TextureVisuals = TextureVisuals

    def load(path, force=None):
        ply = load_ply(path)
        m = types.SimpleNamespace(vertices=ply.vertices, faces=ply.faces,
                                  vertex_normals=ply.vertex_normals if ply.vertex_normals is not None else np.zeros_like(ply.vertices))
        if ply.uv is not None and ply.texture_image is not None:
            m.visual = TextureVisuals()
            m.visual.uv = ply.uv.copy()
            m.visual.material = types.SimpleNamespace(image=np.asarray(ply.texture_image)[:, :, :3])
        else:
            m.visual = types.SimpleNamespace(vertex_colors=ply.vertex_colors)
        return m

    tm.load = load
    sys.modules["trimesh"], sys.modules["trimesh.visual"], sys.modules["trimesh.visual.texture"] = tm, tm.visual, tm.visual.texture

    # --- pyrr: only Quaternion(list) is reached ------------------------------------------------------------------
    py = types.ModuleType("pyrr")
    py.Quaternion = lambda r: np.asarray(r, dtype=np.float64)
    sys.modules["pyrr"] = py
    torch.nn.Module.cuda = lambda self, *a, **k: self


from cube_scenario import CUBE_C, CUBE_CAM, CUBE_F, CUBE_Q, CUBE_T, CUBE_V, cube_targets  # noqa: E402


def run_untextured(ref, cfg_base):
    """Second scenario: the vertex-colour branch (diffdope.py:229-231,1677-1686) on a cube, default loss config
    (mask only) AND all three, 3 hypotheses x 3 iterations."""
    import tempfile

    import cv2
    from omegaconf import OmegaConf

    tmp = tempfile.mkdtemp()
    ply = os.path.join(tmp, "cube.ply")
    with open(ply, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 8\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\n"
                "property uchar green\nproperty uchar blue\nelement face 12\nproperty list uchar int vertex_indices\nend_header\n")
        for v, c in zip(CUBE_V, CUBE_C):
            f.write("%.9g %.9g %.9g %d %d %d\n" % (v[0], v[1], v[2], c[0], c[1], c[2]))
        for t in CUBE_F:
            f.write("3 %d %d %d\n" % tuple(t))
    rgb, depth, seg = cube_targets()
    cv2.imwrite(os.path.join(tmp, "rgb.png"), rgb[..., ::-1])
    cv2.imwrite(os.path.join(tmp, "depth.png"), depth)
    cv2.imwrite(os.path.join(tmp, "seg.png"), seg)
    out = {}
    for tag, all_losses in (("cube_mask", False), ("cube_all", True)):
        cfg = OmegaConf.create(OmegaConf.to_container(cfg_base))
        cfg.camera = CUBE_CAM
        cfg.scene.path_img, cfg.scene.path_depth, cfg.scene.path_segmentation = [os.path.join(tmp, n) for n in ("rgb.png", "depth.png", "seg.png")]
        cfg.scene.image_resize = 1.0
        cfg.losses.l1_rgb_with_mask = all_losses
        cfg.losses.l1_depth_with_mask = all_losses
        cfg.losses.l1_mask = True
        cfg.hyperparameters.batchsize = 3
        cfg.hyperparameters.nb_iterations = 2
        cfg.hyperparameters.learning_rates_bound = [0.05, 0.5]
        obj = ref.Object3D(position=list(CUBE_T / 0.01), rotation=list(CUBE_Q), batchsize=3, opencv2opengl=False, model_path=ply, scale=0.01)
        random.seed(1)
        d = ref.DiffDope(cfg=cfg, object3d=obj)
        out[tag + "_lr"] = d.learning_rates.numpy().copy()
        d.run_optimization()
        out[tag + "_final"] = np.stack([getattr(obj, n).detach().numpy().copy() for n in ("qx", "qy", "qz", "qw", "x", "y", "z")], 1)
        out[tag + "_keys"] = np.array(list(d.losses_values.keys()))
        for k, v in d.losses_values.items():
            out[tag + "_loss_" + k] = v.numpy()
        out[tag + "_rgb0_sample"] = d.optimization_results[0]["rgb"][:, ::5, ::7].numpy()
        out[tag + "_argmin"] = np.array(int(d.get_argmin()))
        print(tag, "keys", list(d.losses_values.keys()), "final[0]", out[tag + "_final"][0])
    return out


def main():
    from omegaconf import OmegaConf  # the repo's stand-in (compat/)

    q, t = su.example_pose()  # the config pose in the OpenGL convention (uses the repo's package: before it is shadowed)
    install_shims()
    # `import diffdope as dd` inside the reference must not pick up the repo's package: give it a bare module
    for k in [k for k in sys.modules if k == "diffdope" or k.startswith("diffdope.")]:
        del sys.modules[k]
    dd = types.ModuleType("diffdope")
    dd.xfm_points = lambda points, matrix, use_python=False: refpath.xfm_points(points, matrix)
    sys.modules["diffdope"] = dd
    ref = mrv.import_reference()
    dd.l1_rgb_with_mask, dd.l1_depth_with_mask, dd.l1_mask = ref.l1_rgb_with_mask, ref.l1_depth_with_mask, ref.l1_mask

    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
    cfg.scene.image_resize = RESIZE
    for k in ("path_img", "path_depth", "path_segmentation"):
        cfg.scene[k] = os.path.join(ROOT, cfg.scene[k])
    cfg.losses.l1_rgb_with_mask = True
    cfg.losses.l1_depth_with_mask = True
    cfg.losses.l1_mask = True
    cfg.hyperparameters.batchsize = B
    cfg.hyperparameters.nb_iterations = NB_ITER
    cfg.hyperparameters.learning_rates_bound = [0.05, 0.5]  # the non-expanding regime
    scale = float(cfg.object3d.scale)
    obj = ref.Object3D(position=list(np.asarray(t, np.float64) / scale), rotation=list(np.asarray(q, np.float64)), batchsize=B,
                       opencv2opengl=False, model_path=os.path.join(ROOT, cfg.object3d.model_path), scale=scale)
    random.seed(0)
    ddope = ref.DiffDope(cfg=cfg, object3d=obj)
    lr = ddope.learning_rates.numpy().copy()
    pose0 = np.stack([getattr(obj, n).detach().numpy().copy() for n in ("qx", "qy", "qz", "qw", "x", "y", "z")], 1)
    ddope.run_optimization()
    final = np.stack([getattr(obj, n).detach().numpy().copy() for n in ("qx", "qy", "qz", "qw", "x", "y", "z")], 1)
    out = dict(resize=RESIZE, lr=lr, pose0=pose0, final=final,
               mtx=np.stack([r["mtx"].numpy() for r in ddope.optimization_results]),
               rgb_sum=np.array([float(r["rgb"].double().sum()) for r in ddope.optimization_results]),
               depth_sum=np.array([float(r["depth"].double().sum()) for r in ddope.optimization_results]),
               rgb0_sample=ddope.optimization_results[0]["rgb"][:, ::7, ::9].numpy(),
               depth0_sample=ddope.optimization_results[0]["depth"][:, ::7, ::9].numpy(),
               mask_last_sample=ddope.optimization_results[-1]["mask"][:, ::7, ::9].numpy(),
               argmin=np.array(int(ddope.get_argmin())), best_pose=np.asarray(ddope.get_pose()),
               loss_keys=np.array(list(ddope.losses_values.keys())),
               **{"loss_" + k: v.numpy() for k, v in ddope.losses_values.items()})
    out.update(run_untextured(ref, cfg))
    np.savez_compressed(os.path.join(HERE, "reference_run.npz"), **out)
    print("wrote reference_run.npz; keys", list(ddope.losses_values.keys()), "argmin", int(ddope.get_argmin()))
    print("losses first/last:", {k: (v[0].numpy(), v[-1].numpy()) for k, v in ddope.losses_values.items()})


if __name__ == "__main__":
    main()
