"""Shared scene setup for tests, smoke() and bench.py: the reference's example scene
(`configs/diffdope.yaml`, `data/example`) loaded the way the reference loads it
(`diffdope/diffdope.py:784-851,1122-1152`)."""
import os
import random

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data", "example")

CAMERA = dict(fx=1390.53, fy=1386.99, cx=964.957, cy=522.586, im_width=1920, im_height=1080)
POSITION = [-161.16877980209404, 206.22094040904116, 747.151333695172]
ROTATION = [-0.7913458966114294, 0.07584660081839613, 0.6066456668109877, 0.46529349746608056, 0.7183778584745024,
            0.5171413865369608, -0.39657739866517305, 0.6915059982370961, -0.6037763006860087]
SCALE = 0.01


def load_image(path, resize=1.0, depth=False, depth_scale=100):
    """`Image.__post_init__` (`diffdope/diffdope.py:1122-1152`) -> float32 numpy."""
    if depth:
        im = cv2.imread(path, cv2.IMREAD_UNCHANGED) / depth_scale
    else:
        im = cv2.imread(path)[:, :, :3]
        im = cv2.cvtColor(im, cv2.COLOR_BGR2RGB) / 255.0
    im = cv2.flip(im, 0)
    if resize < 1.0:
        size = (int(im.shape[1] * resize), int(im.shape[0] * resize))
        im = cv2.resize(im, size, interpolation=cv2.INTER_NEAREST) if depth else cv2.resize(im, size)
    return np.ascontiguousarray(im, dtype=np.float32)


def example_mesh_arrays():
    from diffdope._ply import load_ply

    m = load_ply(os.path.join(DATA, "mesh", "AlphabetSoup.ply"))
    uv = m.uv.copy()
    uv[:, 1] = 1 - uv[:, 1]
    return dict(
        pos=(m.vertices.astype(np.float32) * np.float32(SCALE)).astype(np.float32),
        tri=m.faces.astype(np.int32),
        uv=uv.astype(np.float32),
        tex=(m.texture_image / 255.0).astype(np.float32),
    )


def example_pose():
    from diffdope._quat import opencv_2_opengl, rotation_to_quat

    t, q = opencv_2_opengl(np.array(POSITION) * SCALE, rotation_to_quat(ROTATION))
    return q.astype(np.float32), t.astype(np.float32)


def example_targets(resize):
    sc = os.path.join(DATA, "scene")
    return dict(
        rgb=load_image(os.path.join(sc, "rgb.png"), resize),
        depth=load_image(os.path.join(sc, "depth.png"), resize, depth=True),
        segmentation=load_image(os.path.join(sc, "seg.png"), resize),
    )


def projection():
    from oracle.refpath import projection_matrix

    c = CAMERA
    return projection_matrix(c["fx"], c["fy"], c["cx"], c["cy"], c["im_width"], c["im_height"])


def projection_native():
    """Same matrix without importing the oracle (product-side code path)."""
    c = CAMERA
    w, h, zn, zf = c["im_width"], c["im_height"], 0.01, 200.0
    d = float(zf - zn)
    return np.array([[2 * c["fx"] / w, 0, (-2 * c["cx"] + w) / w, 0], [0, 2 * c["fy"] / h, (2 * c["cy"] - h) / h, 0],
                     [0, 0, -(zf + zn) / d, -2 * zf * zn / d], [0, 0, -1, 0]], dtype=np.float64)


def lr_multipliers(B, lo=0.01, hi=100, seed=0):
    """`DiffDope.set_batchsize` draws (`diffdope/diffdope.py:1368-1374`) with a seed."""
    random.seed(seed)
    return np.array([random.uniform(lo, hi) for _ in range(B)], dtype=np.float32)


def centred_window(seg, size, H, W):
    """size x size window centred on the bbox centre of seg > 0, clamped to the frame."""
    ys, xs = np.nonzero(seg[..., 0] > 0)
    cy, cx = (ys.min() + ys.max()) // 2, (xs.min() + xs.max()) // 2
    h, w = min(size, H), min(size, W)
    y0 = int(min(max(cy - h // 2, 0), H - h))
    x0 = int(min(max(cx - w // 2, 0), W - w))
    return (y0, x0, h, w)


def perturbed_poses(q, t, B, seed=1, rot_deg=3.0, trans=0.03):
    """Distinct starting hypotheses around (q,t) (superset of the reference, which starts all equal)."""
    rng = np.random.default_rng(seed)
    qs = np.tile(q, (B, 1)).astype(np.float64)
    ts = np.tile(t, (B, 1)).astype(np.float64)
    qs[1:] += rng.normal(0, np.deg2rad(rot_deg) / 2, size=(B - 1, 4))
    ts[1:] += rng.normal(0, trans, size=(B - 1, 3))
    return qs.astype(np.float32), ts.astype(np.float32)


def record(name, **meas):
    """Measured parity figures -> gpurun_out/parity_measured.jsonl (when that directory exists), quoted in DESIGN.md."""
    import json

    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: np.asarray(v, dtype=np.float64).tolist() for k, v in meas.items()}}) + "\n")


def grad_rel_err(go, gg):
    return float(np.abs(np.asarray(go) - np.asarray(gg)).max() / np.abs(np.asarray(go)).max())
