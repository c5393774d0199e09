"""GPU checks of the host-side edges added around the hot path: integer image samples converted on the device
(`Image.set_raw`, extension) and the empty hypothesis shard of a rank that has nothing to refine."""
import os

import cv2
import numpy as np
import pytest
import torch

import scene_util as su

pytestmark = pytest.mark.gpu


def test_image_set_raw_on_the_device_equals_the_host_pipeline():
    """`Image.set_raw` (one kernel, `ddope_image_from_raw`) against `Image.__post_init__` (cv2 on the host): full size and the
    default config's img_resize = 0.5, with and without the flip, 3- and 4-channel colour, uint8 and uint16 depth."""
    import diffdope as dd

    sc = os.path.join(su.DATA, "scene")
    for name, depth in (("rgb.png", False), ("seg.png", False), ("depth.png", True)):
        path = os.path.join(sc, name)
        raw = cv2.imread(path, cv2.IMREAD_UNCHANGED if depth else cv2.IMREAD_COLOR)
        raw = raw.view(np.int16) if raw.dtype == np.uint16 else raw
        pinned = torch.from_numpy(np.ascontiguousarray(raw)).pin_memory()
        for resize in (1.0, 0.5):
            host = dd.Image(img_path=path, depth=depth, img_resize=resize).img_tensor
            dev = dd.Image(depth=depth, img_resize=resize).set_raw(pinned, device="cuda").img_tensor
            assert dev.is_cuda and dev.dtype == torch.float32 and dev.is_contiguous()
            assert torch.equal(dev.cpu(), host), (name, resize)
        noflip = dd.Image(depth=depth, flip_img=False).set_raw(raw).img_tensor
        assert torch.equal(noflip.cpu(), torch.flip(dd.Image(img_path=path, depth=depth).img_tensor, dims=[0]))
    bgra = np.concatenate([cv2.imread(os.path.join(sc, "rgb.png")), np.full((1080, 1920, 1), 255, np.uint8)], -1)
    assert torch.equal(dd.Image().set_raw(bgra).img_tensor.cpu(), dd.Image(img_path=os.path.join(sc, "rgb.png")).img_tensor)
    d8 = (cv2.imread(os.path.join(sc, "depth.png"), cv2.IMREAD_UNCHANGED) >> 4).astype(np.uint8)
    assert torch.equal(dd.Image(depth=True, img_resize=0.5).set_raw(d8).img_tensor.cpu(),
                       torch.tensor(cv2.resize(cv2.flip(d8 / 100, 0), (960, 540), interpolation=cv2.INTER_NEAREST)).float())
    with pytest.raises(ValueError):
        dd.Image(img_resize=0.3).set_raw(cv2.imread(os.path.join(sc, "rgb.png")))
    with pytest.raises(ValueError):
        dd.Image(img_resize=0.5).set_raw(np.zeros((11, 10, 3), np.uint8))


def test_empty_shard_enqueues_nothing():
    from diffdope import _native as nat

    arr = su.example_mesh_arrays()
    gt = su.example_targets(0.25)
    H, W = gt["rgb"].shape[:2]
    sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(su.projection(), H, W)
    g = {k: torch.from_numpy(v).cuda() for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"])
    q = torch.empty(0, 4, device="cuda")
    t = torch.empty(0, 3, device="cuda")
    lr = torch.empty(0, device="cuda")
    ph, lh = sc.optimize(q, t, lr, [1.0, 0.5, 0.25], nat.make_loss_cfg(True, True, True), b_global=4)
    assert tuple(ph.shape) == (3, 0, 7) and tuple(lh.shape) == (3, 0, nat.NUM_LOSSES)


def test_three_channel_segmentation_with_differing_channels_matches_oracle():
    """The reference multiplies per channel by a [H,W,3] segmentation (`diffdope.py:552-555,598`); its example's
    channels are identical, so this case needs its own target: channel 1 switched off in the left half, channel 2 at
    half weight."""
    from oracle import refpath
    from test_gpu_parity import ALL, Example, _cfg, _loss_table

    ex = Example(0.5)
    seg = ex.gt["segmentation"].copy()
    seg[:, : ex.W // 2, 1] = 0.0
    seg[..., 2] *= 0.5
    ex.gt["segmentation"] = seg
    ex.sc.set_target(ex.g["rgb"], ex.g["depth"], torch.from_numpy(seg).cuda())
    B = 2
    qs, ts = su.perturbed_poses(ex.q, ex.t, B)
    lr = su.lr_multipliers(B)
    loss, grad = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, ALL))
    logged, gq, gtr, _ = refpath.forward_backward(ex.oracle_mesh(), ex.P, qs, ts, ex.gt_t(), lr, ALL, ex.H, ex.W)
    assert np.allclose(loss.cpu().numpy(), _loss_table(logged, B), rtol=1e-4, atol=1e-10)
    go, gg = np.concatenate([gq, gtr], 1), grad.cpu().numpy()
    assert np.abs(go - gg).max() <= 1e-4 * np.abs(go).max()
    # and it is a different problem from the identical-channel one
    ex.sc.set_target(ex.g["rgb"], ex.g["depth"], ex.g["segmentation"])
    loss_same, _ = ex.sc.loss_grad(torch.from_numpy(qs).cuda(), torch.from_numpy(ts).cuda(), torch.from_numpy(lr).cuda(), _cfg(ex.n, ALL))
    assert not np.allclose(loss.cpu().numpy()[:, [0, 2]], loss_same.cpu().numpy()[:, [0, 2]], rtol=1e-3)
