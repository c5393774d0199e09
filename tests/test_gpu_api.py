"""GPU (-m gpu): the reference-facing Python API end to end (DiffDope / Scene / Object3D / Camera,
Hydra-style config), the autograd path for user-written losses, and the example script."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

import scene_util as su

pytestmark = pytest.mark.gpu
ROOT = su.ROOT


def _cfg(**over):
    import diffdope  # noqa: F401
    from omegaconf import OmegaConf

    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
    for k in ("path_img", "path_depth", "path_segmentation"):
        cfg.scene[k] = os.path.join(ROOT, cfg.scene[k])
    cfg.object3d.model_path = os.path.join(ROOT, cfg.object3d.model_path)
    for k, v in over.items():
        OmegaConf.update(cfg, k, v)
    return cfg


def test_diffdope_default_config_matches_oracle():
    """configs/diffdope.yaml as shipped (mask loss only, B=8 -> here 3, 960x540), 10 iterations,
    learning-rate multipliers seeded like SURVEY.md 8(d): per-iteration losses and final pose vs the oracle."""
    import diffdope as dd
    from oracle import refpath

    cfg = _cfg(**{"hyperparameters.batchsize": 3, "hyperparameters.nb_iterations": 9})
    random.seed(0)
    ddope = dd.DiffDope(cfg=cfg)
    assert np.allclose(ddope.learning_rates.cpu().numpy(), su.lr_multipliers(3))
    ddope.run_optimization()
    assert list(ddope.losses_values.keys()) == ["mask_selection"] and tuple(ddope.losses_values["mask_selection"].shape) == (10, 3)
    assert len(ddope.optimization_results) == 10
    arr, (q, t), gt = su.example_mesh_arrays(), su.example_pose(), su.example_targets(0.5)
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    hyper = dict(nb_iterations=9, base_lr=20, lr_decay=0.1, learning_rate_base=1)
    o = refpath.run_optimization(mesh, su.projection(), np.tile(q, (3, 1)), np.tile(t, (3, 1)), {k: torch.from_numpy(v) for k, v in gt.items()},
                                 su.lr_multipliers(3), dict(l1_mask=True, weight_mask=1), hyper, 540, 960)
    ours, ref = ddope.losses_values["mask_selection"].numpy(), o["losses"]["mask_selection"]
    # the multipliers are 84x / 76x / 42x: late iterations of the large-step hypotheses are chaotic (the
    # reference itself is not run-to-run reproducible there: unordered atomics),
    # so the trajectory is compared tightly while it is well conditioned and by final pose after that
    assert np.allclose(ours[:6], ref[:6], rtol=5e-4)
    assert np.allclose(ours, ref, rtol=0.06)
    assert int(ddope.get_argmin()) == refpath.argmin_hypothesis(o["losses"])
    best = int(ddope.get_argmin())
    assert np.allclose(ddope.get_pose(), o["mtx"][-1][best], atol=1e-3)
    qf, tf = ddope.object3d.pose_tensors()
    assert np.abs(tf.cpu().numpy()[best] - o["final"][best, 4:]).max() < 1e-3  # 0.1 mm
    qo = o["final"][best, :4]
    qg = qf.cpu().numpy()[best]
    cosang = abs(float((qo / np.linalg.norm(qo)) @ (qg / np.linalg.norm(qg))))
    assert np.degrees(2 * np.arccos(min(cosang, 1.0))) < 0.1
    # lazily re-rendered history entry equals a direct render of that iteration's pose
    res = ddope.optimization_results[-1]
    assert tuple(res["rgb"].shape) == (3, 540, 960, 3) and tuple(res["depth"].shape) == (3, 540, 960) and tuple(res["mask"].shape) == (3, 540, 960, 3)
    img = ddope.render_img()
    assert img.ndim == 3 and img.shape[2] == 3 and img.dtype == np.uint8
    single = ddope.render_img(index=3, batch_index=1)
    assert single.ndim == 3
    plot = ddope.plot_losses()
    assert plot.shape == (600, 1000, 3)


def test_user_loss_through_autograd_matches_fused_path():
    """A hand-written torch loss equal to the built-in stack, appended as a custom function, goes
    through render_texture_batch's CUDA backward and must give the fused path's trajectory."""
    import diffdope as dd

    cfg = _cfg(**{"hyperparameters.batchsize": 2, "hyperparameters.nb_iterations": 5, "losses.l1_rgb_with_mask": True, "losses.l1_depth_with_mask": True,
                  "hyperparameters.learning_rates_bound": [0.1, 0.3]})  # non-chaotic regime, see test_gpu_parity
    random.seed(0)
    a = dd.DiffDope(cfg=cfg)
    a.run_optimization()
    random.seed(0)
    b = dd.DiffDope(cfg=cfg)

    def my_loss(d):
        seg = d.gt_tensors["segmentation"]
        l = dd.dist_batch_lr(torch.abs((d.renders["rgb"] - d.gt_tensors["rgb"]) * seg), d.learning_rates).mean() * 0.7
        l = l + dd.dist_batch_lr(torch.abs((d.renders["depth"] - d.gt_tensors["depth"]) * seg[..., 0]), d.learning_rates, [1, 2]).mean()
        diff = torch.abs(d.renders["mask"] - seg)
        d.add_loss_value("mine", torch.mean(diff.detach(), (1, 2, 3)))
        return l + dd.dist_batch_lr(diff, d.learning_rates).mean()

    b.loss_functions = [my_loss]
    b.run_optimization()
    qa, ta = a.object3d.pose_tensors()
    qb, tb = b.object3d.pose_tensors()
    assert np.abs((ta - tb).cpu().numpy()).max() < 1e-3 and np.abs((qa - qb).cpu().numpy()).max() < 1e-3
    assert np.allclose(b.losses_values["mine"].numpy(), a.losses_values["mask_selection"].numpy(), rtol=2e-3)
    assert isinstance(b.optimization_results[-1], dict) and tuple(b.optimization_results[-1]["rgb"].shape) == (2, 540, 960, 3)


def test_render_texture_batch_gradient_matches_oracle():
    import diffdope as dd
    from oracle import refpath

    arr, (q, t), gt = su.example_mesh_arrays(), su.example_pose(), su.example_targets(0.25)
    H, W = gt["rgb"].shape[:2]
    qs, ts = su.perturbed_poses(q, t, 2)
    # feed the canonical matrices so both sides rasterise exactly the same geometry
    from oracle import nvdr

    _, M = nvdr.canonical_pose(qs, ts)
    mtx = torch.from_numpy(M).cuda().requires_grad_(True)
    dev = "cuda"
    rr = dd.render_texture_batch(None, torch.from_numpy(su.projection().astype(np.float32)).to(dev), mtx, torch.from_numpy(arr["pos"]).to(dev),
                                 torch.from_numpy(arr["tri"]).to(dev), [H, W], uv=torch.from_numpy(arr["uv"]).to(dev), tex=torch.from_numpy(arr["tex"]).to(dev))
    g = {k: torch.from_numpy(v).to(dev) for k, v in gt.items()}
    wr, wd, wm = torch.rand(2, H, W, 3, device=dev), torch.rand(2, H, W, device=dev), torch.rand(2, H, W, 3, device=dev)
    ((rr["rgb"] * wr).sum() + (rr["depth"] * wd).sum() + (rr["mask"] * wm).sum()).backward()
    mt = torch.from_numpy(M).requires_grad_(True)
    mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    # oracle with an explicit matrix: reuse its graph below the pose
    proj_t = torch.from_numpy(su.projection().astype(np.float32))
    mvp = refpath._with_value(torch.matmul(proj_t.expand(2, 4, 4), mt), nvdr.canonical_mvp(su.projection(), M))
    pos = torch.from_numpy(mesh.pos)
    pos_clip = refpath.xfm_points(pos[None].expand(2, -1, -1), mvp)
    rast = refpath._Rasterize.apply(pos_clip, mesh.tri, H, W)
    gb = refpath._Interpolate.apply(torch.cat([pos, torch.ones(pos.shape[0], 1)], 1), rast, mesh.tri)
    depth = refpath.xfm_points(gb.reshape(2, -1, 4)[..., :3], mt).reshape(2, H, W, 4)[..., 2] * -1
    mask = refpath._Antialias.apply(refpath._Interpolate.apply(torch.ones(pos.shape[0], 3), rast, mesh.tri), rast, pos_clip, mesh.tri, mesh.opp)
    color = refpath._TextureLinear.apply(torch.from_numpy(mesh.tex), refpath._Interpolate.apply(torch.from_numpy(mesh.uv), rast, mesh.tri))
    color = color * torch.clamp(rast[..., -1:], 0, 1)
    ((color * wr.cpu()).sum() + (depth * wd.cpu()).sum() + (mask * wm.cpu()).sum()).backward()
    go, gg = mt.grad.numpy(), mtx.grad.cpu().numpy()
    su.record("render_texture_batch_autograd_bridge", grad_rel_err=su.grad_rel_err(go, gg))
    assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
    assert np.array_equal(rr["rgb"].detach().cpu().numpy(), color.detach().numpy())


def test_example_script_runs(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "diff-dope_b200"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "simple_scene.py"), "hyperparameters.nb_iterations=6", "hyperparameters.batchsize=4",
                          "hydra.run.dir=%s" % tmp_path], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "best hypothesis" in out.stdout
    for f in ("plot.png", "simple_scene.mp4"):
        p = os.path.join(ROOT, f)
        assert os.path.exists(p) and os.path.getsize(p) > 1000
        os.remove(p)


def test_bop_style_object_loop(tmp_path):
    """examples/run_bop_scene.py: the reference's per-object flow (run_bop_scene.py:48-93) -- Mesh built
    and batched separately, Object3D without a model path, segmentation Image swapped into the scene."""
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "diff-dope_b200"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_bop_scene.py"), "hyperparameters.nb_iterations=4", "hyperparameters.batchsize=3",
                          "hydra.run.dir=%s" % tmp_path], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "object 0: best hypothesis" in out.stdout
    assert os.path.getsize(os.path.join(tmp_path, "00.png")) > 1000


def test_batched_bop_example_runs(tmp_path):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "diff-dope_b200"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_bop_scene_batched.py"), "hyperparameters.nb_iterations=4",
                          "hyperparameters.batchsize=3", "hydra.run.dir=%s" % tmp_path], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "object 2: best hypothesis" in out.stdout
    assert os.path.getsize(os.path.join(tmp_path, "02.png")) > 1000


def test_reference_example_runs_unchanged(tmp_path):
    """The reference's own `examples/simple_scene.py` with the reference's own `configs/diffdope.yaml`, both byte-for-byte
    copies shipped as test inputs under tests/golden/ (tests/test_host.py checks them against /root/reference where that
    exists), run unmodified against this repository's package: default config (960x540, 8 hypotheses, mask loss, 61
    iterations), `ic(get_argmin(), get_pose())`, plot.png, simple_scene.mp4."""
    script = os.path.join(ROOT, "tests", "golden", "reference_examples", "simple_scene.py")
    # the reference script imports hydra before diffdope: the stand-ins must be importable up front
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "diff-dope_b200", "compat")]))
    for f in ("plot.png", "simple_scene.mp4"):
        if os.path.exists(os.path.join(ROOT, f)):
            os.remove(os.path.join(ROOT, f))
    out = subprocess.run([sys.executable, script, "hydra.run.dir=%s" % tmp_path], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Saved animation to simple_scene.mp4" in out.stdout
    assert "ic| " in out.stdout + out.stderr  # ic(ddope.get_argmin(), ddope.get_pose())
    for f in ("plot.png", "simple_scene.mp4"):
        p = os.path.join(ROOT, f)
        assert os.path.exists(p) and os.path.getsize(p) > 1000
        os.remove(p)


def test_reference_bop_example_runs_unchanged(tmp_path):
    """The reference's own `examples/run_bop_scene.py` (byte-for-byte copy under tests/golden/reference_examples/) against this
    package. Its BOP paths are hard-coded to its author's home directory (run_bop_scene.py:19-25), so the script runs under
    tests/run_with_path_map.py, which serves those prefixes from a BOP-layout directory built here: the reference's own perturbed-pose
    file for HOPE val/000001 (18 objects in frame "0"; copy under tests/golden/), the example rgb / depth images, the example
    segmentation as every object's mask_visib, and the example mesh under every obj_id. The script itself is not touched."""
    import json
    import shutil

    scene = tmp_path / "hope" / "val" / "000001"
    models = tmp_path / "hope" / "models"
    for d in (scene / "rgb", scene / "depth", scene / "mask_visib", models, tmp_path / "data" / "hope" / "val" / "000001", tmp_path / "out"):
        d.mkdir(parents=True)
    ex = os.path.join(ROOT, "data", "example")
    shutil.copy(os.path.join(ex, "scene", "rgb.png"), scene / "rgb" / "000000.png")
    shutil.copy(os.path.join(ex, "scene", "depth.png"), scene / "depth" / "000000.png")
    poses = os.path.join(ROOT, "tests", "golden", "hope_val_000001_scene_error_deg_040_trans_016.json")
    shutil.copy(poses, tmp_path / "data" / "hope" / "val" / "000001" / "scene_error_deg_040_trans_016.json")
    frame0 = json.load(open(poses))["0"]
    assert len(frame0) == 18
    for k, obj in enumerate(frame0):
        os.symlink(os.path.join(ex, "scene", "seg.png"), scene / "mask_visib" / ("000000_%06d.png" % k))
        ply = models / ("obj_%06d.ply" % obj["obj_id"])
        if not ply.exists():
            os.symlink(os.path.join(ex, "mesh", "AlphabetSoup.ply"), ply)
    for f in os.listdir(os.path.join(ex, "mesh")):  # the PLY names its texture file: it must sit next to the model
        if not f.endswith(".ply"):
            os.symlink(os.path.join(ex, "mesh", f), models / f)
    path_map = [["/home/jtremblay/code/camera2robot/hope", str(tmp_path / "hope")], ["/home/jtremblay/code/diff-dope/data", str(tmp_path / "data")]]
    env = dict(os.environ, PATH_MAP=json.dumps(path_map),
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "diff-dope_b200", "compat")]))
    script = os.path.join(ROOT, "tests", "golden", "reference_examples", "run_bop_scene.py")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_with_path_map.py"), script, "hyperparameters.nb_iterations=6",
                          "hyperparameters.batchsize=4", "hydra.run.dir=%s" % (tmp_path / "out")], capture_output=True, text=True, cwd=ROOT, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    text = out.stdout + out.stderr
    assert "object 17" in text, "all 18 objects of the frame were refined"
    for k in range(18):
        assert os.path.getsize(tmp_path / "out" / ("%02d.png" % k)) > 1000


def test_batched_objects_equal_sequential_loop():
    """dd.run_optimization_batched (all objects of a frame as one batch of launches) gives each object
    bit for bit what the reference-style sequential loop gives; two of the objects share one Mesh."""
    import diffdope as dd

    cfg = _cfg(**{"hyperparameters.batchsize": 4, "hyperparameters.nb_iterations": 5, "losses.l1_rgb_with_mask": True,
                  "losses.l1_depth_with_mask": True})
    random.seed(0)
    base = dd.DiffDope(cfg=cfg)
    mesh = base.object3d.mesh
    other = dd.Mesh(cfg.object3d.model_path, scale=cfg.object3d.scale * 0.8)
    other.set_batchsize(4)
    other.cuda()
    pos = np.array(cfg.object3d.position, dtype=np.float64)

    def make(k):
        obj = dd.Object3D(position=list(pos + np.array([3.0 * k, -2.0 * k, 4.0 * k])), rotation=cfg.object3d.rotation, scale=cfg.object3d.scale, batchsize=4)
        obj.mesh = other if k == 1 else mesh
        obj.cuda()
        d = dd.DiffDope(cfg=cfg, camera=base.camera, object3d=obj, scene=base.scene)
        d.learning_rates = base.learning_rates.clone() * (1.0 + 0.1 * k)
        return d

    seq = [make(k) for k in range(3)]
    for d in seq:
        d.run_optimization()
    # one launch over (object, hypothesis) with per-object mesh / target tables (ddope_optimize_multi), and one stream per object
    for one_launch in (True, False):
        par = dd.run_optimization_batched([make(k) for k in range(3)], one_launch=one_launch)
        for a, b in zip(seq, par):
            assert list(a.losses_values.keys()) == list(b.losses_values.keys()) == ["rgb", "depth", "mask_selection"]
            for k in a.losses_values:
                assert torch.equal(a.losses_values[k], b.losses_values[k]), (one_launch, k)
            qa, ta = a.object3d.pose_tensors()
            qb, tb = b.object3d.pose_tensors()
            assert torch.equal(qa, qb) and torch.equal(ta, tb)
            assert np.array_equal(a.get_pose(), b.get_pose())
            assert torch.equal(a._pose_hist_host, b._pose_hist_host)
    assert not torch.equal(seq[0].losses_values["rgb"], seq[1].losses_values["rgb"])


def test_extension_config_keys_through_the_api():
    """losses.l1_edge, hyperparameters.optimizer=adam and render.texture_filter reach the CUDA path from the
    Hydra-style config; the fused edge loss equals the torch-written l1_edge run through autograd."""
    import diffdope as dd

    over = {"hyperparameters.batchsize": 2, "hyperparameters.nb_iterations": 2, "losses.l1_mask": False, "losses.l1_edge": True,
            "losses.weight_edge": 0.5}
    # moderate multipliers: with the default draws (84x, 76x) the sign-gradient iteration amplifies 1e-5 differences
    # between the two gradient paths into percent-level loss differences within two steps ("chaotic")
    small = torch.tensor([0.3, 1.0]).cuda()
    random.seed(0)
    a = dd.DiffDope(cfg=_cfg(**over))
    a.learning_rates = small.clone()
    a.run_optimization()
    assert list(a.losses_values.keys()) == ["edge"] and tuple(a.losses_values["edge"].shape) == (3, 2)
    assert not torch.equal(a.losses_values["edge"][0], a.losses_values["edge"][2]), "the pose must move"
    random.seed(0)
    b = dd.DiffDope(cfg=_cfg(**over))
    b.learning_rates = small.clone()

    def my_edge(d):  # not in the fused table -> generic autograd path
        return dd.l1_edge(d)

    b.loss_functions = [my_edge]
    b.run_optimization()
    assert np.allclose(a.losses_values["edge"].numpy(), b.losses_values["edge"].numpy(), rtol=2e-3, atol=1e-8)
    qa, ta = a.object3d.pose_tensors()
    qb, tb = b.object3d.pose_tensors()
    assert (qa - qb).abs().max() < 1e-4 and (ta - tb).abs().max() < 1e-4
    # Adam + mipmaps from the config
    random.seed(0)
    c = dd.DiffDope(cfg=_cfg(**{"hyperparameters.batchsize": 2, "hyperparameters.nb_iterations": 3, "hyperparameters.optimizer": "adam",
                                "hyperparameters.base_lr": 0.01, "render.texture_filter": "linear-mipmap-linear", "losses.l1_rgb_with_mask": True}))
    c.run_optimization()
    q0, t0 = c._pose_hist_host[0, :, :4].cpu(), c._pose_hist_host[0, :, 4:].cpu()
    q1, t1 = c._pose_hist_host[1, :, :4].cpu(), c._pose_hist_host[1, :, 4:].cpu()
    lr0 = 0.01 * 0.1 ** 1.0
    assert np.allclose((q1 - q0).abs().numpy(), lr0, rtol=5e-3) and np.allclose((t1 - t0).abs().numpy(), lr0, rtol=5e-3)


def test_api_matches_the_reference_loop_fixture():
    """The product's DiffDope API on the GPU against tests/golden/reference_run.npz: the output of the reference's own,
    unmodified run_optimization executed on the CPU (nvdiffrast ops served by the oracle; make_reference_run.py).
    Same config, same seeded learning-rate draws, same start pose."""
    import diffdope as dd

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz"))
    cfg = _cfg(**{"scene.image_resize": float(g["resize"]), "hyperparameters.batchsize": 2, "hyperparameters.nb_iterations": 3,
                  "hyperparameters.learning_rates_bound": [0.05, 0.5], "losses.l1_rgb_with_mask": True, "losses.l1_depth_with_mask": True,
                  "losses.l1_mask": True})
    random.seed(0)
    d = dd.DiffDope(cfg=cfg)
    assert np.allclose(d.learning_rates.cpu().numpy(), g["lr"])
    q0, t0 = d.object3d.pose_tensors()
    assert np.allclose(torch.cat([q0, t0], 1).cpu().numpy(), g["pose0"], atol=1e-6)
    d.run_optimization()
    assert list(d.losses_values.keys()) == list(g["loss_keys"])
    H, W = d.resolution
    for k in d.losses_values:
        ours, ref = d.losses_values[k].numpy(), g["loss_" + k]
        assert ours.shape == ref.shape
        assert np.allclose(ours, ref, rtol=5e-4, atol=2.0 / (H * W)), k
    qf, tf = d.object3d.pose_tensors()
    assert np.abs(torch.cat([qf, tf], 1).cpu().numpy() - g["final"]).max() < 1e-4
    assert int(d.get_argmin()) == int(g["argmin"])
    assert np.abs(d.get_pose() - g["best_pose"]).max() < 1e-4
    for i in range(4):
        assert np.abs(d.optimization_results[i]["mtx"].numpy() - g["mtx"][i]).max() < 1e-4
    res0 = d.optimization_results[0]
    assert np.allclose(res0["rgb"].numpy()[:, ::7, ::9], g["rgb0_sample"], atol=1e-6)
    assert np.allclose(res0["depth"].numpy()[:, ::7, ::9], g["depth0_sample"], atol=1e-5)
    assert np.allclose(d.optimization_results[-1]["mask"].numpy()[:, ::7, ::9], g["mask_last_sample"], atol=1e-4)


def test_api_matches_the_reference_loop_fixture_untextured(tmp_path):
    """Vertex-colour branch: the product's API on the GPU (own PLY reader, own image loader, untextured kernels) against
    the reference's unmodified loop run on the CPU (tests/golden/reference_run.npz, scenario cube_*)."""
    import importlib.util

    import cv2
    import diffdope as dd
    from omegaconf import OmegaConf

    spec = importlib.util.spec_from_file_location("cube_scenario", os.path.join(os.path.dirname(__file__), "golden", "cube_scenario.py"))
    cs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cs)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz"))
    ply = str(tmp_path / "cube.ply")
    with open(ply, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 8\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\n"
                "property uchar green\nproperty uchar blue\nelement face 12\nproperty list uchar int vertex_indices\nend_header\n")
        for v, c in zip(cs.CUBE_V, cs.CUBE_C):
            f.write("%.9g %.9g %.9g %d %d %d\n" % (v[0], v[1], v[2], c[0], c[1], c[2]))
        for t in cs.CUBE_F:
            f.write("3 %d %d %d\n" % tuple(t))
    rgb, depth, seg = cs.cube_targets()
    cv2.imwrite(str(tmp_path / "rgb.png"), rgb[..., ::-1])
    cv2.imwrite(str(tmp_path / "depth.png"), depth)
    cv2.imwrite(str(tmp_path / "seg.png"), seg)
    for tag, all_losses in (("cube_mask", False), ("cube_all", True)):
        cfg = _cfg(**{"scene.image_resize": 1.0, "hyperparameters.batchsize": 3, "hyperparameters.nb_iterations": 2,
                      "hyperparameters.learning_rates_bound": [0.05, 0.5], "losses.l1_rgb_with_mask": all_losses,
                      "losses.l1_depth_with_mask": all_losses, "losses.l1_mask": True})
        cfg.camera = OmegaConf.create(dict(cs.CUBE_CAM))
        cfg.scene.path_img, cfg.scene.path_depth, cfg.scene.path_segmentation = str(tmp_path / "rgb.png"), str(tmp_path / "depth.png"), str(tmp_path / "seg.png")
        obj = dd.Object3D(position=list(cs.CUBE_T / 0.01), rotation=list(cs.CUBE_Q), batchsize=3, opencv2opengl=False, model_path=ply, scale=0.01)
        random.seed(1)
        d = dd.DiffDope(cfg=cfg, object3d=obj)
        assert np.allclose(d.learning_rates.cpu().numpy(), g[tag + "_lr"])
        d.run_optimization()
        assert list(d.losses_values.keys()) == list(g[tag + "_keys"])
        for k in d.losses_values:
            assert np.allclose(d.losses_values[k].numpy(), g[tag + "_loss_" + k], rtol=5e-4, atol=2.0 / (72 * 96)), (tag, k)
        qf, tf = d.object3d.pose_tensors()
        assert np.abs(torch.cat([qf, tf], 1).cpu().numpy() - g[tag + "_final"]).max() < 2e-4
        assert int(d.get_argmin()) == int(g[tag + "_argmin"])
        assert np.allclose(d.optimization_results[0]["rgb"].numpy()[:, ::5, ::7], g[tag + "_rgb0_sample"], atol=1e-6)
