"""The five BASELINE.json configs as concrete, seeded workloads. Shared by the
parity tests and bench.py. Configs 3-5 are STAND-INS: BOP YCB-V / T-LESS models and images are not in the
tree, so meshes are procedural (or rescaled copies of the example mesh) and targets are renders of the
ground-truth pose -- every report must label them as such.

Nothing here imports the oracle or the CUDA library: a workload is plain numpy data plus a description;
the caller renders the targets (oracle at small sizes, the CUDA renderer at full size)."""
import json
import os

import numpy as np

import scene_util as su

F = np.float32


def uv_sphere(stacks, slices, radius=0.5, flatten=(1.0, 1.0, 1.0)):
    """Latitude/longitude sphere without seam duplication: V = (stacks-1)*slices + 2,
    T = 2*slices*(stacks-1). uv = (longitude/2pi, latitude/pi)."""
    v, uv = [(0.0, 0.0, 1.0)], [(0.5, 0.0)]
    for i in range(1, stacks):
        th = np.pi * i / stacks
        for j in range(slices):
            ph = 2 * np.pi * j / slices
            v.append((np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)))
            uv.append((j / slices, i / stacks))
    v.append((0.0, 0.0, -1.0))
    uv.append((0.5, 1.0))
    ring = lambda i, j: 1 + (i - 1) * slices + (j % slices)
    f = []
    for j in range(slices):
        f.append((0, ring(1, j), ring(1, j + 1)))
    for i in range(1, stacks - 1):
        for j in range(slices):
            a, b, c, d = ring(i, j), ring(i + 1, j), ring(i + 1, j + 1), ring(i, j + 1)
            f.append((a, b, c))
            f.append((a, c, d))
    last = len(v) - 1
    for j in range(slices):
        f.append((last, ring(stacks - 1, j + 1), ring(stacks - 1, j)))
    pos = np.array(v, dtype=np.float64) * radius * np.array(flatten)
    return pos.astype(F), np.array(f, dtype=np.int32), np.array(uv, dtype=F)


def procedural_texture(size=2048, seed=0):
    """`numpy.random.default_rng(0).random((size,size,3), float32)` box-blurred 5x5 (wrap)."""
    t = np.random.default_rng(seed).random((size, size, 3), dtype=F)
    acc = np.zeros_like(t)
    for dy in range(-2, 3):
        for dx in range(-2, 3):
            acc += np.roll(np.roll(t, dy, 0), dx, 1)
    return (acc / F(25.0)).astype(F)


def projection(fx, fy, cx, cy, w, h, zn=0.01, zf=200.0):
    """`Camera.get_projection_matrix` (`diffdope/diffdope.py:679-742`)."""
    d = float(zf - zn)
    return np.array([[2 * fx / w, 0, (-2 * cx + w) / w, 0], [0, 2 * fy / h, (2 * cy - h) / h, 0],
                     [0, 0, -(zf + zn) / d, -2 * zf * zn / d], [0, 0, -1, 0]], dtype=np.float64)


def axis_angle_quat(axis, deg):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    h = np.deg2rad(deg) / 2
    return np.array([a[0] * np.sin(h), a[1] * np.sin(h), a[2] * np.sin(h), np.cos(h)])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def config4(scale=1.0):
    """T-LESS stand-in: textureless 10,000-triangle flattened sphere, 720x540, depth + mask losses."""
    pos, tri, _ = uv_sphere(51, 100, 0.5, (1.0, 1.0, 0.6))
    W, H = int(720 * scale), int(540 * scale)
    P = projection(1075 * scale, 1075 * scale, 360 * scale, 270 * scale, W, H)
    q_gt, t_gt = np.array([0.0, 0.0, 0.0, 1.0]), np.array([0.0, 0.0, -7.0])
    q0 = quat_mul(axis_angle_quat((1, 1, 0), 10.0), q_gt)
    t0 = t_gt + 0.04 * np.array([1.0, -1.0, 1.0]) / np.sqrt(3.0)
    return dict(name="config4: T-LESS stand-in (procedural 10k-triangle textureless sphere, oracle/CUDA-rendered target)",
                pos=pos, tri=tri, uv=None, tex=None, vtx_color=np.full((pos.shape[0], 3), 0.5, F), P=P, H=H, W=W,
                q_gt=q_gt.astype(F), t_gt=t_gt.astype(F), q0=q0.astype(F), t0=t0.astype(F), B=256, iters=100,
                losses=dict(l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0), window=None)


def config5(scale=1.0, tex_size=2048):
    """Synthetic stress: textured 50,000-triangle sphere filling ~50 % of a 1024^2 window, full loss stack."""
    pos, tri, uv = uv_sphere(126, 200, 1.0, (1.0, 1.0, 0.8))
    W = H = int(1024 * scale)
    P = projection(1400 * scale, 1400 * scale, 512 * scale, 512 * scale, W, H)
    q_gt = axis_angle_quat((0.3, 1.0, 0.2), 35.0)
    t_gt = np.array([0.0, 0.0, -3.43])
    q0 = quat_mul(axis_angle_quat((1, -1, 0.5), 4.0), q_gt)
    t0 = t_gt + np.array([0.02, -0.015, 0.03])
    return dict(name="config5: synthetic stress (procedural 50k-triangle textured sphere, 2048^2 procedural texture)",
                pos=pos, tri=tri, uv=uv, tex=procedural_texture(tex_size), vtx_color=None, P=P, H=H, W=W,
                q_gt=q_gt.astype(F), t_gt=t_gt.astype(F), q0=q0.astype(F), t0=t0.astype(F), B=1024, iters=50,
                losses=dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0,
                            l1_edge=True, weight_edge=0.5), window=None)


def config3(scale=1.0, n_objects=8):
    """BOP YCB-V stand-in: 8 objects (the example mesh rescaled by 0.6..1.3) at the poses of YCB-V scene 000048
    frame 1 (cycled to reach 8), YCB-V-like intrinsics at 640x480, full loss stack incl. Sobel edge."""
    from diffdope._quat import opencv_2_opengl, rotation_to_quat

    fix = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ycbv_000048_frame1.json")))
    arr = su.example_mesh_arrays()
    W, H = int(640 * scale), int(480 * scale)
    P = projection(1066.778 * scale, 1067.487 * scale, 312.9869 * scale, 241.3109 * scale, W, H)
    factors = np.linspace(0.6, 1.3, n_objects)
    objs = []
    for k in range(n_objects):
        g, i = fix["gt"][k % len(fix["gt"])], fix["init"][k % len(fix["init"])]
        # cycled copies are shifted sideways so the eight objects do not coincide
        shift = np.array([60.0 * (k // len(fix["gt"])), 40.0 * (k // len(fix["gt"])), 0.0])
        tg, qg = opencv_2_opengl((np.array(g["cam_t_m2c"]) + shift) * 0.01, rotation_to_quat(g["cam_R_m2c"]))
        ti, qi = opencv_2_opengl((np.array(i["cam_t_m2c"]) + shift) * 0.01, rotation_to_quat(i["cam_R_m2c"]))
        objs.append(dict(pos=(arr["pos"] * F(factors[k])).astype(F), tri=arr["tri"], uv=arr["uv"], tex=arr["tex"], vtx_color=None,
                         q_gt=qg.astype(F), t_gt=tg.astype(F), q0=qi.astype(F), t0=ti.astype(F), obj_id=g["obj_id"]))
    return dict(name="config3: BOP YCB-V stand-in (8 rescaled copies of the example mesh at YCB-V 000048/frame-1 poses, rendered targets)",
                objects=objs, P=P, H=H, W=W, B=128, iters=100,
                losses=dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0,
                            l1_edge=True, weight_edge=0.5), window=None)


def targets_from_render(rgb, depth, rast_id):
    """Target images of a stand-in workload from a render of the ground-truth pose: rgb as rendered,
    depth where covered (0 elsewhere, like a depth sensor's holes), segmentation = coverage in 3 channels."""
    cov = (rast_id > 0).astype(F)
    return dict(rgb=np.ascontiguousarray(rgb, dtype=F), depth=np.ascontiguousarray(depth * cov, dtype=F),
                segmentation=np.ascontiguousarray(np.repeat(cov[..., None], 3, -1), dtype=F))
