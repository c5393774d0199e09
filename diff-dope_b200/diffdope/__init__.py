"""B200-native Diff-DOPE: same public names as NVlabs/diff-dope's `diffdope` package
(`diffdope/__init__.py:1-7`), hot path in libddope_b200.so (hand-written sm_100a CUDA)."""
from . import _compat

_compat.ensure()

from .diffdope import *  # noqa: F401,F403,E402
from .diffdope import (Camera, DiffDope, Image, Mesh, Object3D, Scene, dist_batch_lr, find_crop, getimg_stack, interpolate,  # noqa: F401
                       l1_depth_with_mask, l1_mask, l1_rgb_with_mask, make_grid, make_grid_image, make_grid_overlay_batch,
                       matrix_batch_44_from_position_quat, opencv_2_opengl, render_texture_batch,
                       l1_edge, run_optimization_batched, sobel_magnitude)  # the last three are extensions
from .ops import xfm_points, xfm_vectors  # noqa: F401

__all__ = ["xfm_points", "xfm_vectors"]
