"""Quaternion / rotation helpers in the conventions the reference gets from pyrr.

The reference uses `pyrr.Quaternion`, `pyrr.Matrix33(...).quaternion`,
`pyrr.Matrix44(...).quaternion`, `Quaternion.matrix44` and the Hamilton product
(`diffdope/diffdope.py:103,124,128-137,1002-1004`). pyrr is not a dependency
here; these functions restate the conventions (pyrr 0.10.x, UPSTREAM):
quaternions are `[x, y, z, w]`, matrices act on column vectors, `Matrix33(list9)`
is row-major.
"""
import numpy as np


def quat_from_matrix(m):
    """Trace-based rotation-matrix -> quaternion (x,y,z,w), column-vector convention."""
    m = np.asarray(m, dtype=np.float64)
    if m.size == 9:
        m = m.reshape(3, 3)
    m = m[:3, :3]
    trace = m[0, 0] + m[1, 1] + m[2, 2]
    if trace > 0:
        s = 0.5 / np.sqrt(trace + 1.0)
        qx = (m[2, 1] - m[1, 2]) * s
        qy = (m[0, 2] - m[2, 0]) * s
        qz = (m[1, 0] - m[0, 1]) * s
        qw = 0.25 / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = 2.0 * np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2])
        qx = 0.25 * s
        qy = (m[0, 1] + m[1, 0]) / s
        qz = (m[0, 2] + m[2, 0]) / s
        qw = (m[2, 1] - m[1, 2]) / s
    elif m[1, 1] > m[2, 2]:
        s = 2.0 * np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2])
        qx = (m[0, 1] + m[1, 0]) / s
        qy = 0.25 * s
        qz = (m[1, 2] + m[2, 1]) / s
        qw = (m[0, 2] - m[2, 0]) / s
    else:
        s = 2.0 * np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1])
        qx = (m[0, 2] + m[2, 0]) / s
        qy = (m[1, 2] + m[2, 1]) / s
        qz = 0.25 * s
        qw = (m[1, 0] - m[0, 1]) / s
    return np.array([qx, qy, qz, qw], dtype=np.float64)


def quat_to_matrix33(q):
    """Quaternion (x,y,z,w, any norm) -> 3x3 rotation acting on column vectors."""
    qx, qy, qz, qw = [float(v) for v in q]
    sqx, sqy, sqz, sqw = qx * qx, qy * qy, qz * qz, qw * qw
    invs = 1.0 / (sqx + sqy + sqz + sqw)
    m = np.empty((3, 3), dtype=np.float64)
    m[0, 0] = (sqx - sqy - sqz + sqw) * invs
    m[1, 1] = (-sqx + sqy - sqz + sqw) * invs
    m[2, 2] = (-sqx - sqy + sqz + sqw) * invs
    m[1, 0] = 2.0 * (qx * qy + qz * qw) * invs
    m[0, 1] = 2.0 * (qx * qy - qz * qw) * invs
    m[2, 0] = 2.0 * (qx * qz - qy * qw) * invs
    m[0, 2] = 2.0 * (qx * qz + qy * qw) * invs
    m[2, 1] = 2.0 * (qy * qz + qx * qw) * invs
    m[1, 2] = 2.0 * (qy * qz - qx * qw) * invs
    return m


def quat_mul(a, b):
    """Hamilton product a*b for (x,y,z,w) quaternions."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            ax * bw + ay * bz - az * by + aw * bx,
            -ax * bz + ay * bw + az * bx + aw * by,
            ax * by - ay * bx + az * bw + aw * bz,
            -ax * bx - ay * by - az * bz + aw * bw,
        ],
        dtype=np.float64,
    )


def quat_axis(axis, theta):
    h = 0.5 * theta
    q = np.zeros(4, dtype=np.float64)
    q["xyz".index(axis)] = np.sin(h)
    q[3] = np.cos(h)
    return q


def rotation_to_quat(rotation):
    """Reference `Object3D.set_pose` input rule (`diffdope/diffdope.py:1000-1004`):
    4 values -> quaternion (x,y,z,w); 9 values or 3x3 -> row-major matrix."""
    r = np.asarray(rotation, dtype=np.float64)
    n = len(rotation)
    assert n == 4 or n == 3 or n == 9
    if n == 4:
        return r.reshape(4).copy()
    return quat_from_matrix(r.reshape(3, 3))


def opencv_2_opengl(p, q):
    """OpenCV -> OpenGL camera-frame change of a pose (p, q).

    Restates `diffdope/diffdope.py:92-140`: R' = diag(1,-1,-1) R,
    t' = diag(1,-1,-1) t, followed by the reference's "legacy" quaternion block
    (four axis rotations whose product is the identity rotation; kept so the
    quaternion sign matches the reference's).
    """
    flip = np.diag([1.0, -1.0, -1.0])
    rot = flip @ quat_to_matrix33(q)
    t = flip @ np.asarray(p, dtype=np.float64).reshape(3)
    q2 = quat_from_matrix(rot)
    q2 = quat_mul(quat_mul(q2, quat_axis("z", np.pi / 2)), quat_axis("y", -np.pi / 2))
    q2 = quat_mul(quat_mul(q2, quat_axis("z", -np.pi / 2)), quat_axis("x", -np.pi / 2))
    return t, q2
