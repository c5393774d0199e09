"""ctypes binding of libddope_b200.so (include/ddope_b200.h) for torch tensors.

The library is plain C ABI: this module passes `tensor.data_ptr()` and the current
CUDA stream. There is no CPU fallback: if the shared library is missing or a tensor
is not on a CUDA device the call raises.
"""
import ctypes
import os

import numpy as np
import torch

_LIB = None
# DDOPE_B200_LIB: alternative build of the same library (A/B timing of compile-time variants); never a fallback
_LIB_PATH = os.environ.get("DDOPE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libddope_b200.so")

NUM_LOSSES = 4
LOSS_KEYS = ("rgb", "depth", "mask_selection", "edge")  # reference add_loss_value keys, diffdope.py:559,577,605; "edge" is an extension
OPT_SGD, OPT_ADAM = 0, 1
TEX_LINEAR, TEX_MIPMAP = 0, 1


class LossCfg(ctypes.Structure):
    _fields_ = [
        ("use_rgb", ctypes.c_int32),
        ("use_depth", ctypes.c_int32),
        ("use_mask", ctypes.c_int32),
        ("weight_rgb", ctypes.c_float),
        ("weight_depth", ctypes.c_float),
        ("weight_mask", ctypes.c_float),
        ("use_edge", ctypes.c_int32),
        ("weight_edge", ctypes.c_float),
    ]


class OptimCfg(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("beta1", ctypes.c_float),
        ("beta2", ctypes.c_float),
        ("eps", ctypes.c_float),
        ("step0", ctypes.c_int32),
    ]


def lib_path():
    return _LIB_PATH


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            "libddope_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C diff-dope_b200/csrc` (there is no CPU / PyTorch fallback for the hot path)" % _LIB_PATH
        )
    L = ctypes.CDLL(_LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.ddope_abi_version.restype = ci
    L.ddope_last_error.restype = ctypes.c_char_p
    L.ddope_last_launch_count.restype = ctypes.c_int64
    L.ddope_last_launch_count.argtypes = [vp]
    L.ddope_graph_launch_count.restype = ctypes.c_int64
    L.ddope_graph_launch_count.argtypes = [vp]
    L.ddope_scene_set_graph.argtypes = [vp, ci]
    L.ddope_xfm_fwd.argtypes = [vp, ci, ci, vp, ci, ci, vp, vp]
    L.ddope_xfm_bwd.argtypes = [vp, ci, ci, vp, ci, vp, vp]
    L.ddope_xfm_bwd_mtx.argtypes = [vp, ci, ci, vp, ci, ci, vp, vp]
    L.ddope_xfm_bwd_full.argtypes = [vp, ci, ci, vp, vp, ci, ci, vp, vp, vp]
    L.ddope_scene_create.argtypes = [ctypes.POINTER(vp), vp, ci, vp, ci, vp, vp, ci, ci, vp]
    L.ddope_scene_destroy.argtypes = [vp]
    L.ddope_scene_set_camera.argtypes = [vp, vp, ci, ci]
    L.ddope_scene_set_target.argtypes = [vp, vp, vp, vp, ci, vp]
    L.ddope_scene_set_window.argtypes = [vp, ci, ci, ci, ci]
    L.ddope_scene_set_texture_filter.argtypes = [vp, ci, ci]
    L.ddope_scene_set_culling.argtypes = [vp, ci]
    L.ddope_scene_set_raster_mode.argtypes = [vp, ci]
    L.ddope_scene_raster_mode.argtypes = [vp]
    L.ddope_scene_raster_mode.restype = ci
    L.ddope_scene_set_bin_capacity.argtypes = [vp, ci]
    L.ddope_scene_mesh_orientation.argtypes = [vp]
    L.ddope_scene_mesh_orientation.restype = ci
    L.ddope_scene_set_optimizer.argtypes = [vp, ctypes.POINTER(OptimCfg)]
    L.ddope_render.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp, vp]
    L.ddope_render_mtx.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp]
    L.ddope_render_bwd.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp]
    L.ddope_render_bwd_attr.argtypes = [vp, vp, ci, vp, vp, vp, vp]
    L.ddope_scene_update_texture.argtypes = [vp, vp, vp]
    L.ddope_scene_update_vertex_colors.argtypes = [vp, vp, vp]
    L.ddope_image_from_raw.argtypes = [vp, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, vp, vp]
    L.ddope_image_from_raw.restype = ci
    L.ddope_debug_read.argtypes = [vp, ci, vp, ctypes.c_int64]
    L.ddope_debug_read.restype = ctypes.c_int64
    L.ddope_profile_begin.argtypes = [vp]
    L.ddope_profile_end.argtypes = [vp, vp, vp]
    L.ddope_loss_grad.argtypes = [vp, vp, vp, vp, ci, ci, ctypes.POINTER(LossCfg), vp, vp, vp]
    L.ddope_optimize.argtypes = [vp, vp, vp, vp, ci, ci, vp, ci, ctypes.POINTER(LossCfg), vp, vp, vp]
    L.ddope_optimize_multi.argtypes = [vp, ci, vp, vp, vp, vp, vp, ci, vp, ci, ctypes.POINTER(LossCfg), vp, vp, vp]
    for name in (
        "ddope_xfm_fwd", "ddope_xfm_bwd", "ddope_xfm_bwd_mtx", "ddope_xfm_bwd_full", "ddope_scene_create",
        "ddope_scene_destroy", "ddope_scene_set_camera", "ddope_scene_set_target", "ddope_scene_set_window",
        "ddope_render", "ddope_render_mtx", "ddope_render_bwd", "ddope_loss_grad", "ddope_optimize",
        "ddope_profile_begin", "ddope_profile_end", "ddope_scene_set_texture_filter", "ddope_scene_set_optimizer",
        "ddope_scene_set_culling", "ddope_image_from_raw", "ddope_scene_set_raster_mode", "ddope_scene_set_bin_capacity", "ddope_scene_set_graph", "ddope_optimize_multi", "ddope_render_bwd_attr",
        "ddope_scene_update_texture", "ddope_scene_update_vertex_colors",
    ):
        getattr(L, name).restype = ci
    if L.ddope_abi_version() != 2:
        raise RuntimeError("libddope_b200.so ABI version mismatch")
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError("ddope_b200: " + lib().ddope_last_error().decode("utf-8", "replace"))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a cuda tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _host(a, dtype):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


def _hptr(a):
    return ctypes.c_void_p(0 if a is None else a.ctypes.data)


def make_loss_cfg(use_rgb, use_depth, use_mask, w_rgb=1.0, w_depth=1.0, w_mask=1.0, use_edge=False, w_edge=1.0):
    return LossCfg(int(bool(use_rgb)), int(bool(use_depth)), int(bool(use_mask)), float(w_rgb), float(w_depth), float(w_mask),
                   int(bool(use_edge)), float(w_edge))


def image_from_raw(raw, is_depth, divisor, flip=True, resize_half=False):
    """`ddope_image_from_raw`: raw cuda uint8 / uint16 (int16 view) samples [H,W,C] or [H,W] -> float32 [h,w,3] RGB or [h,w] depth,
    bit-equal to the reference's host pipeline (`Image.__post_init__`, diffdope.py:1122-1152)."""
    if not isinstance(raw, torch.Tensor) or not raw.is_cuda:
        raise RuntimeError("raw samples must be a cuda tensor")
    raw = raw.contiguous()
    nbytes = raw.element_size()
    if raw.dtype not in (torch.uint8, torch.uint16, torch.int16):
        raise RuntimeError("raw samples must be uint8 or uint16")
    H, W = int(raw.shape[0]), int(raw.shape[1])
    C = 1 if raw.dim() == 2 else int(raw.shape[2])
    oh, ow = (H // 2, W // 2) if resize_half else (H, W)
    out = torch.empty((oh, ow) if (is_depth or C == 1) else (oh, ow, 3), dtype=torch.float32, device=raw.device)
    _check(lib().ddope_image_from_raw(_ptr(raw), nbytes, H, W, C, int(bool(is_depth)), float(divisor), int(bool(flip)), int(bool(resize_half)),
                                      _ptr(out), _stream()))
    return out


def optimize_multi(scenes, counts, b_globals, quat, trans, lr_mult, lr_sched, cfg, out=None):
    """`ddope_optimize_multi`: the hypotheses of several objects (scenes[k] has counts[k] of them, concatenated in that order in
    quat [B,4] / trans [B,3] / lr_mult [B]; b_globals[k] = divisor of object k's hypothesis mean) refined by one sequence of
    launches. In place on quat / trans; returns (pose_hist [n,B,7], loss_hist [n,B,4])."""
    for t, n in ((quat, "quat"), (trans, "trans")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("%s must be a contiguous cuda float32 tensor (updated in place)" % n)
    lr_mult = _dev_f32(lr_mult, "lr_mult")
    B = int(quat.shape[0])
    if B != sum(int(c) for c in counts) or len(counts) != len(scenes) or len(b_globals) != len(scenes):
        raise RuntimeError("optimize_multi: counts / b_globals do not match the scenes and the hypothesis arrays")
    sched = np.ascontiguousarray(lr_sched, dtype=np.float32)
    n = sched.shape[0]
    if out is not None:
        pose_hist, loss_hist = out
    else:
        pose_hist = torch.empty(n, B, 7, device=quat.device)
        loss_hist = torch.empty(n, B, NUM_LOSSES, device=quat.device)
    if B == 0:
        return pose_hist, loss_hist
    hyp_scene = np.ascontiguousarray(np.repeat(np.arange(len(scenes), dtype=np.int32), np.asarray(counts, dtype=np.int64)))
    hyp_bg = np.ascontiguousarray(np.repeat(np.asarray(b_globals, dtype=np.int32), np.asarray(counts, dtype=np.int64)))
    handles = (ctypes.c_void_p * len(scenes))(*[sc._h for sc in scenes])
    _check(lib().ddope_optimize_multi(ctypes.cast(handles, ctypes.c_void_p), len(scenes), _hptr(hyp_scene), _hptr(hyp_bg), _ptr(quat), _ptr(trans),
                                      _ptr(lr_mult), B, _hptr(sched), n, ctypes.byref(cfg), _ptr(pose_hist), _ptr(loss_hist), _stream()))
    return pose_hist, loss_hist


class NativeScene:
    """Owns one `ddope_scene`: a mesh uploaded once plus camera, target, window."""

    def __init__(self, pos, tri, uv=None, tex=None, vtx_color=None):
        L = lib()
        pos = _host(pos, np.float32).reshape(-1, 3)
        tri = _host(tri, np.int32).reshape(-1, 3)
        uv = None if uv is None else _host(uv, np.float32).reshape(-1, 2)
        tex = None if tex is None else _host(tex, np.float32)
        if tex is not None:
            if tex.ndim != 3 or tex.shape[2] < 3:
                raise RuntimeError("tex must be [H,W,3]")
            tex = np.ascontiguousarray(tex[:, :, :3])
        vc = None if vtx_color is None else _host(vtx_color, np.float32).reshape(-1, 3)
        self.V, self.T = pos.shape[0], tri.shape[0]
        self.textured = uv is not None and tex is not None
        h = ctypes.c_void_p()
        th, tw = (tex.shape[0], tex.shape[1]) if self.textured else (0, 0)
        _check(L.ddope_scene_create(ctypes.byref(h), _hptr(pos), self.V, _hptr(tri), self.T,
                                    _hptr(uv if self.textured else None), _hptr(tex if self.textured else None), th, tw,
                                    _hptr(None if self.textured else vc)))
        self._h = h
        self._keep = {}
        self.tex_shape = tuple(tex.shape[:2]) if self.textured else None
        self.H = self.W = None
        self.window = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ddope_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_camera(self, proj, H, W):
        p = _host(proj, np.float32).reshape(4, 4)
        _check(lib().ddope_scene_set_camera(self._h, _hptr(p), int(H), int(W)))
        self.H, self.W = int(H), int(W)
        self.window = (0, 0, self.H, self.W)
        self._keep.pop("target", None)

    def set_window(self, y0, x0, h, w):
        _check(lib().ddope_scene_set_window(self._h, int(y0), int(x0), int(h), int(w)))
        self.window = (int(y0), int(x0), int(h), int(w))

    def set_texture_filter(self, mode="linear", max_levels=0):
        """'linear' (reference, default) or 'linear-mipmap-linear' (extension)."""
        modes = {"linear": TEX_LINEAR, "linear-mipmap-linear": TEX_MIPMAP, TEX_LINEAR: TEX_LINEAR, TEX_MIPMAP: TEX_MIPMAP}
        if mode not in modes:
            raise RuntimeError("ddope_b200: unknown texture filter %r" % (mode,))
        _check(lib().ddope_scene_set_texture_filter(self._h, modes[mode], int(max_levels)))

    def set_culling(self, auto=True):
        """Back-face culling of closed, consistently oriented meshes (default on); False rasterises every triangle."""
        _check(lib().ddope_scene_set_culling(self._h, 1 if auto else 0))

    def set_raster_mode(self, mode, bin_capacity=None):
        """'zbuffer' (one launch over all triangles, global 64-bit atomicMin z-buffer) or 'binned' (per-tile triangle bins staged by
        TMA, shared-memory z-buffer inside the pixel pass). Same raster rule, bit-identical results."""
        modes = {"zbuffer": 0, "binned": 1, 0: 0, 1: 1}
        if mode not in modes:
            raise RuntimeError("ddope_b200: unknown raster mode %r" % (mode,))
        _check(lib().ddope_scene_set_raster_mode(self._h, modes[mode]))
        if bin_capacity is not None:
            _check(lib().ddope_scene_set_bin_capacity(self._h, int(bin_capacity)))

    def raster_mode(self):
        return ("zbuffer", "binned")[int(lib().ddope_scene_raster_mode(self._h))]

    def bin_overflows(self):
        """Tile bins that overflowed (and fell back to scanning the mesh) since the scene was created."""
        return int(self.debug_read(2, 4).view(np.int32)[0])

    def mesh_orientation(self):
        """+1 / -1: closed mesh of positive / negative volume (culling applies); 0: open or inconsistent mesh."""
        return int(lib().ddope_scene_mesh_orientation(self._h))

    def set_optimizer(self, kind="sgd", beta1=0.9, beta2=0.999, eps=1e-8, step0=0):
        """'sgd' (reference, default) or 'adam' (extension; torch.optim.Adam's algebra)."""
        kinds = {"sgd": OPT_SGD, "adam": OPT_ADAM}
        if str(kind).lower() not in kinds:
            raise RuntimeError("ddope_b200: unknown optimizer %r" % (kind,))
        cfg = OptimCfg(kinds[str(kind).lower()], float(beta1), float(beta2), float(eps), int(step0))
        _check(lib().ddope_scene_set_optimizer(self._h, ctypes.byref(cfg)))

    def set_target(self, rgb=None, depth=None, seg=None):
        """rgb [H,W,3], depth [H,W], seg [H,W,3] or [H,W] or [H,W,1]; cuda float32, borrowed."""
        seg_c = 0
        if rgb is not None:
            rgb = _dev_f32(rgb, "rgb")
            assert tuple(rgb.shape) == (self.H, self.W, 3), "rgb target must be [H,W,3]"
        if depth is not None:
            depth = _dev_f32(depth, "depth")
            assert tuple(depth.shape) == (self.H, self.W), "depth target must be [H,W]"
        if seg is not None:
            seg = _dev_f32(seg, "seg")
            if seg.dim() == 2:
                seg_c = 1
            else:
                seg_c = seg.shape[2]
            assert tuple(seg.shape[:2]) == (self.H, self.W) and seg_c in (1, 3), "seg target must be [H,W] or [H,W,1|3]"
        self._keep["target"] = (rgb, depth, seg)
        _check(lib().ddope_scene_set_target(self._h, _ptr(rgb), _ptr(depth), _ptr(seg), seg_c, _stream()))

    def render(self, quat, trans, want=("rgb", "depth", "mask", "rast", "mtx")):
        quat, trans = _dev_f32(quat, "quat"), _dev_f32(trans, "trans")
        B = quat.shape[0]
        if self.window is None:
            raise RuntimeError("ddope_b200: set the camera first")
        _, _, h, w = self.window
        dev = quat.device
        out = {}
        if "rgb" in want:
            out["rgb"] = torch.empty(B, h, w, 3, device=dev)
        if "depth" in want:
            out["depth"] = torch.empty(B, h, w, device=dev)
        if "mask" in want:
            out["mask"] = torch.empty(B, h, w, device=dev)
        if "rast" in want:
            out["rast"] = torch.empty(B, h, w, 4, device=dev)
        if "mtx" in want:
            out["mtx"] = torch.empty(B, 4, 4, device=dev)
        _check(lib().ddope_render(self._h, _ptr(quat), _ptr(trans), B, _ptr(out.get("rgb")), _ptr(out.get("depth")),
                                  _ptr(out.get("mask")), _ptr(out.get("rast")), _ptr(out.get("mtx")), _stream()))
        return out

    def render_mtx(self, mtx, want_rast=True):
        """Render from explicit model matrices [B,4,4]. Returns rgb, depth, mask [B,h,w], rast."""
        mtx = _dev_f32(mtx, "mtx")
        B = mtx.shape[0]
        if self.window is None:
            raise RuntimeError("ddope_b200: set the camera first")
        _, _, h, w = self.window
        dev = mtx.device
        rgb = torch.empty(B, h, w, 3, device=dev)
        depth = torch.empty(B, h, w, device=dev)
        mask = torch.empty(B, h, w, device=dev)
        rast = torch.empty(B, h, w, 4, device=dev) if want_rast else None
        _check(lib().ddope_render_mtx(self._h, _ptr(mtx), B, _ptr(rgb), _ptr(depth), _ptr(mask), _ptr(rast), _stream()))
        return rgb, depth, mask, rast

    def render_bwd(self, mtx, d_rgb=None, d_depth=None, d_mask=None):
        mtx = _dev_f32(mtx, "mtx")
        B = mtx.shape[0]
        d_rgb = None if d_rgb is None else _dev_f32(d_rgb, "d_rgb")
        d_depth = None if d_depth is None else _dev_f32(d_depth, "d_depth")
        d_mask = None if d_mask is None else _dev_f32(d_mask, "d_mask")
        d_mtx = torch.empty(B, 4, 4, device=mtx.device)
        _check(lib().ddope_render_bwd(self._h, _ptr(mtx), B, _ptr(d_rgb), _ptr(d_depth), _ptr(d_mask), _ptr(d_mtx), _stream()))
        return d_mtx

    def render_attr_grad(self, mtx, d_rgb):
        """dL/d rgb [B,h,w,3] -> dL/d tex [Ht,Wt,3] (textured mesh) or dL/d vtx_color [V,3], summed over the batch."""
        mtx, d_rgb = _dev_f32(mtx, "mtx"), _dev_f32(d_rgb, "d_rgb")
        B = mtx.shape[0]
        if self.textured:
            g = torch.empty(self.tex_shape[0], self.tex_shape[1], 3, device=mtx.device)
            _check(lib().ddope_render_bwd_attr(self._h, _ptr(mtx), B, _ptr(d_rgb), _ptr(g), None, _stream()))
        else:
            g = torch.empty(self.V, 3, device=mtx.device)
            _check(lib().ddope_render_bwd_attr(self._h, _ptr(mtx), B, _ptr(d_rgb), None, _ptr(g), _stream()))
        return g

    def update_colors(self, attr):
        """The texture [Ht,Wt,3] / vertex colours [V,3] changed (an optimizer step): refresh the scene's device copy."""
        attr = _dev_f32(attr.detach(), "attr")
        if self.textured:
            assert tuple(attr.shape) == (self.tex_shape[0], self.tex_shape[1], 3)
            _check(lib().ddope_scene_update_texture(self._h, _ptr(attr), _stream()))
        else:
            assert tuple(attr.shape) == (self.V, 3)
            _check(lib().ddope_scene_update_vertex_colors(self._h, _ptr(attr), _stream()))

    def loss_grad(self, quat, trans, lr_mult, cfg, b_global=None):
        quat, trans = _dev_f32(quat, "quat"), _dev_f32(trans, "trans")
        lr_mult = _dev_f32(lr_mult, "lr_mult")
        B = quat.shape[0]
        loss = torch.empty(B, NUM_LOSSES, device=quat.device)
        grad = torch.empty(B, 7, device=quat.device)
        _check(lib().ddope_loss_grad(self._h, _ptr(quat), _ptr(trans), _ptr(lr_mult), B, int(b_global or B),
                                     ctypes.byref(cfg), _ptr(loss), _ptr(grad), _stream()))
        return loss, grad

    def optimize(self, quat, trans, lr_mult, lr_sched, cfg, b_global=None, keep_history=True, out=None):
        """In-place SGD on quat [B,4] / trans [B,3]. Returns (pose_hist [n,B,7], loss_hist [n,B,4]) or (None, None).
        out = (pose_hist, loss_hist): write the history into these preallocated contiguous cuda float32 tensors."""
        for t, n in ((quat, "quat"), (trans, "trans")):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("%s must be a contiguous cuda float32 tensor (updated in place)" % n)
        lr_mult = _dev_f32(lr_mult, "lr_mult")
        B = quat.shape[0]
        sched = np.ascontiguousarray(lr_sched, dtype=np.float32)
        n = sched.shape[0]
        pose_hist = loss_hist = None
        if out is not None:
            pose_hist, loss_hist = out
            if not (tuple(pose_hist.shape) == (n, B, 7) and tuple(loss_hist.shape) == (n, B, NUM_LOSSES) and pose_hist.is_contiguous()
                    and loss_hist.is_contiguous() and pose_hist.dtype == torch.float32 and loss_hist.dtype == torch.float32 and pose_hist.is_cuda and loss_hist.is_cuda):
                raise RuntimeError("optimize(out=...): need contiguous cuda float32 [n,B,7] and [n,B,%d]" % NUM_LOSSES)
        elif keep_history:
            pose_hist = torch.empty(n, B, 7, device=quat.device)
            loss_hist = torch.empty(n, B, NUM_LOSSES, device=quat.device)
        if B == 0:  # an empty shard (more ranks than hypotheses): nothing to enqueue, empty tables
            return pose_hist, loss_hist
        _check(lib().ddope_optimize(self._h, _ptr(quat), _ptr(trans), _ptr(lr_mult), B, int(b_global or B), _hptr(sched), n,
                                    ctypes.byref(cfg), _ptr(pose_hist), _ptr(loss_hist), _stream()))
        return pose_hist, loss_hist

    KERNELS = ("iter_kernel", "raster_kernel", "pixel_kernel")

    def profile_begin(self):
        _check(lib().ddope_profile_begin(self._h))

    def profile_end(self):
        """-> ({kernel: total ms}, {kernel: launches}) for everything enqueued since profile_begin()."""
        ms = (ctypes.c_float * 3)()
        n = (ctypes.c_int * 3)()
        _check(lib().ddope_profile_end(self._h, ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(n, ctypes.c_void_p)))
        return {k: float(ms[i]) for i, k in enumerate(self.KERNELS)}, {k: int(n[i]) for i, k in enumerate(self.KERNELS)}

    def debug_read(self, what, nbytes):
        """Debug hook: raw bytes of an internal work buffer after the last call (0 = tile partial rows, 1 = HypState records)."""
        buf = np.zeros(int(nbytes), dtype=np.uint8)
        n = lib().ddope_debug_read(self._h, int(what), _hptr(buf), int(nbytes))
        if n < 0:
            _check(-1)
        return buf[:n]

    def set_graph(self, on=True):
        """Small-batch CUDA-graph replay of ddope_optimize (default on)."""
        _check(lib().ddope_scene_set_graph(self._h, 1 if on else 0))

    def graph_launch_count(self):
        """ddope_optimize calls served by a CUDA-graph launch so far (small batches)."""
        return int(lib().ddope_graph_launch_count(self._h))

    def last_launch_count(self):
        return int(lib().ddope_last_launch_count(self._h))
