"""Golden vectors from the REFERENCE's own code, generated in the build container (the reference tree does not exist on
the GPU box). The reference module `diffdope/diffdope.py` is imported from /root/reference with inert stand-ins for
the third-party packages that are missing here (nvdiffrast, pyrr, trimesh, hydra, ...); only functions that touch none
of them are called, unmodified, on CPU tensors:

  matrix_batch_44_from_position_quat (diffdope.py:46-89)      -- `.cuda()` inside it is patched to a no-op
  Camera.get_projection_matrix       (diffdope.py:679-742)
  dist_batch_lr, l1_rgb_with_mask, l1_depth_with_mask, l1_mask (diffdope.py:534-613), values, logged values, autograd
  find_crop                          (diffdope.py:242-274)
  the learning-rate schedule expression of run_optimization (diffdope.py:1657-1661) is copied as a formula

    python tests/golden/make_reference_vectors.py   ->  tests/golden/reference_vectors.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/diffdope/diffdope.py"


class _Anything(types.ModuleType):
    """Module stand-in: any attribute is another stand-in, calling it returns a stand-in (decorators pass through)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Anything(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Anything(self.__name__ + "()")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "hydra", "hydra.utils", "imageio", "nvdiffrast", "nvdiffrast.torch", "pyrr",
                 "trimesh", "icecream", "omegaconf", "diffdope"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["icecream"].ic = print
    torch.Tensor.cuda = lambda self, *a, **k: self  # matrix_batch_44_from_position_quat uploads a constant (diffdope.py:85)
    spec = importlib.util.spec_from_file_location("reference_diffdope_module", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Mock:
    """The attributes of DiffDope the loss functions read (diffdope.py:547-613)."""

    def __init__(self, renders, gt, lr, weights):
        self.renders, self.gt_tensors, self.learning_rates = renders, gt, lr
        self.cfg = types.SimpleNamespace(losses=types.SimpleNamespace(weight_rgb=weights[0], weight_depth=weights[1], weight_mask=weights[2]))
        self.logged = {}
        self.optimization_results = [{}]

    def add_loss_value(self, key, values, values_weighted=None):
        self.logged[key] = values.detach().clone()


def main():
    ref = import_reference()
    rng = np.random.default_rng(7)
    out = {}
    # pose -> matrix
    q = rng.normal(size=(5, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    p = rng.normal(size=(5, 3)).astype(np.float32)
    out["pose_q"], out["pose_p"] = q, p
    out["pose_mtx"] = ref.matrix_batch_44_from_position_quat(torch.from_numpy(q), torch.from_numpy(p)).numpy()
    # projection
    cams = np.array([[1390.53, 1386.99, 964.957, 522.586, 1920, 1080], [1066.778, 1067.487, 312.9869, 241.3109, 640, 480]])
    out["cam_params"] = cams
    out["cam_proj"] = np.stack([ref.Camera(fx=c[0], fy=c[1], cx=c[2], cy=c[3], im_width=int(c[4]), im_height=int(c[5])).get_projection_matrix().numpy() for c in cams])
    # losses: B=3 hypotheses of a 12x16 image
    B, H, W = 3, 12, 16
    rgb = torch.tensor(rng.random((B, H, W, 3)).astype(np.float32), requires_grad=True)
    depth = torch.tensor((3 + rng.random((B, H, W))).astype(np.float32), requires_grad=True)
    mask = torch.tensor(rng.random((B, H, W, 3)).astype(np.float32), requires_grad=True)
    gt = {"rgb": torch.tensor(rng.random((B, H, W, 3)).astype(np.float32)), "depth": torch.tensor((3 + rng.random((B, H, W))).astype(np.float32)),
          "segmentation": torch.tensor(np.repeat((rng.random((1, H, W, 1)) > 0.4).astype(np.float32), 3, -1).repeat(B, 0))}
    lr = torch.tensor([0.5, 30.0, 84.4437], dtype=torch.float32)
    weights = (0.7, 1.0, 1.3)
    mock = _Mock({"rgb": rgb, "depth": depth, "mask": mask}, gt, lr, weights)
    l_rgb, l_depth, l_mask = ref.l1_rgb_with_mask(mock), ref.l1_depth_with_mask(mock), ref.l1_mask(mock)
    (l_rgb + l_depth + l_mask).backward()
    out.update(loss_rgb=rgb.detach().numpy(), loss_depth=depth.detach().numpy(), loss_mask=mask.detach().numpy(), loss_gt_rgb=gt["rgb"].numpy(),
               loss_gt_depth=gt["depth"].numpy(), loss_gt_seg=gt["segmentation"].numpy(), loss_lr=lr.numpy(), loss_weights=np.array(weights, np.float32),
               loss_values=np.array([float(l_rgb), float(l_depth), float(l_mask)], np.float32),
               logged_rgb=mock.logged["rgb"].numpy(), logged_depth=mock.logged["depth"].numpy(), logged_mask=mock.logged["mask_selection"].numpy(),
               grad_rgb=rgb.grad.numpy(), grad_depth=depth.grad.numpy(), grad_mask=mask.grad.numpy())
    # find_crop on a synthetic mask
    m = torch.zeros(40, 60, 3)
    m[11:25, 17:41] = 1.0
    out["crop_mask_box"] = np.array([11, 25, 17, 41])
    out["crop"] = np.array(ref.find_crop(m), dtype=np.int64)
    # Image loading (diffdope.py:1122-1152; cv2 is installed here): the example scene at half resolution, the way
    # configs/diffdope.yaml asks for it. Stored: shapes, sums and a strided sample of every image.
    data = os.path.join(os.path.dirname(os.path.dirname(HERE)), "data", "example", "scene")
    for key, fname, kw in (("rgb", "rgb.png", {}), ("depth", "depth.png", {"depth": True}), ("seg", "seg.png", {})):
        im = ref.Image(img_path=os.path.join(data, fname), img_resize=0.5, **kw).img_tensor
        out["img_%s_shape" % key] = np.array(im.shape)
        out["img_%s_sum" % key] = np.array(float(im.double().sum()))
        out["img_%s_sample" % key] = im[::37, ::41].numpy()
    # visualisation helpers (diffdope.py:313-528): grid + overlay + contour of a small synthetic batch
    yy, xx = np.mgrid[0:40, 0:56]
    fg = np.zeros((3, 40, 56, 3), np.float32)
    for k in range(3):
        blob = ((yy - 18 - 2 * k) ** 2 + (xx - 25 - 3 * k) ** 2) < (9 + k) ** 2
        fg[k][blob] = [0.9 - 0.2 * k, 0.3 + 0.2 * k, 0.5]
    bg = rng.random((3, 40, 56, 3)).astype(np.float32)
    out["viz_fg"], out["viz_bg"] = fg, bg
    out["viz_overlay"] = ref.make_grid_overlay_batch(background=torch.from_numpy(bg), foreground=torch.from_numpy(fg), alpha=0.7, row=2, final_width=300,
                                                     add_background=True, add_contour=True, color_countour=[0.46, 0.73, 0], flip_result=True)
    out["viz_overlay_plain"] = ref.make_grid_overlay_batch(background=torch.from_numpy(bg), foreground=torch.from_numpy(fg), alpha=0.5, row=3, final_width=200,
                                                           add_background=False, add_contour=True, flip_result=False)  # add_contour=False crashes in the reference (alpha_img unbound)
    out["viz_grid"] = ref.make_grid(torch.from_numpy(fg).permute(0, 3, 1, 2), nrow=2).numpy()
    # schedule (formula at diffdope.py:1657-1661 with nb_iterations=60, base_lr=20, lr_decay=0.1)
    out["sched"] = np.array([20 * 0.1 ** (it / 60 + 1) for it in range(61)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote reference_vectors.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
