"""Inputs of the untextured scenario shared by make_reference_run.py (generator) and tests/test_oracle_pins.py (checker):
a vertex-coloured cube, a small pinhole camera and synthetic 72x96 target images. Pure numpy, no side effects."""
import numpy as np

CUBE_V = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=np.float64) * 50.0  # PLY units (scale 0.01)
CUBE_F = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
CUBE_C = (np.random.default_rng(3).random((8, 3)) * 255).astype(np.uint8)
CUBE_Q = np.array([0.31, -0.22, 0.12, 0.91]) / np.linalg.norm([0.31, -0.22, 0.12, 0.91])
CUBE_T = np.array([0.12, -0.07, -4.0])
CUBE_HW = (72, 96)
CUBE_CAM = dict(fx=110.0, fy=105.0, cx=47.3, cy=37.1, im_width=96, im_height=72)


def cube_targets():
    """uint8 rgb, uint16 depth (x 1/100 units), uint8 3-channel segmentation, as they are written to PNG files."""
    rng = np.random.default_rng(11)
    H, W = CUBE_HW
    rgb = (rng.random((H, W, 3)) * 255).astype(np.uint8)
    depth = (350 + rng.random((H, W)) * 100).astype(np.uint16)
    yy, xx = np.mgrid[0:H, 0:W]
    seg = (((yy - 36) ** 2 + (xx - 50) ** 2) < 24 ** 2).astype(np.uint8) * 255
    return rgb, depth, np.repeat(seg[..., None], 3, -1)
