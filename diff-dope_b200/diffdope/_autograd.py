"""torch.autograd bridge for `render_texture_batch`: CUDA forward (`ddope_render_mtx`) and CUDA
backward (`ddope_render_bwd`), so user-written loss functions on `renders["rgb"|"depth"|"mask"]`
backpropagate to the pose matrix exactly as they do through nvdiffrast in the reference
(`diffdope/diffdope.py:156-234,1706-1714`)."""
import torch


class _RenderMtx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mtx, scene):
        mtx_c = mtx.detach().contiguous()
        rgb, depth, mask, rast = scene.render_mtx(mtx_c)
        ctx.scene = scene
        ctx.state = (scene.H, scene.W, scene.window)
        ctx.save_for_backward(mtx_c)
        ctx.mark_non_differentiable(rast)
        return rgb, depth, mask, rast

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_mask, _d_rast):
        (mtx,) = ctx.saved_tensors
        sc = ctx.scene
        if (sc.H, sc.W, sc.window) != ctx.state:
            raise RuntimeError("render_texture_batch: the scene's camera/window changed between forward and backward")
        d_mtx = sc.render_bwd(mtx, d_rgb, d_depth, d_mask)
        return d_mtx, None


def render_mtx(scene, mtx):
    return _RenderMtx.apply(mtx, scene)
