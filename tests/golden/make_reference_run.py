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
for p in (ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "diff-dope_b200", "compat"), os.path.join(ROOT, "tests")):
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
    np.savez_compressed(os.path.join(HERE, "reference_run.npz"), **out)
    print("wrote reference_run.npz; keys", list(ddope.losses_values.keys()), "argmin", int(ddope.get_argmin()))
    print("losses first/last:", {k: (v[0].numpy(), v[-1].numpy()) for k, v in ddope.losses_values.items()})


if __name__ == "__main__":
    main()
