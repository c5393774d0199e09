"""torch.autograd bridge for `render_texture_batch`: CUDA forward (`ddope_render_mtx`) and CUDA
backward (`ddope_render_bwd`), so user-written loss functions on `renders["rgb"|"depth"|"mask"]`
backpropagate to the pose matrix exactly as they do through nvdiffrast in the reference
(`diffdope/diffdope.py:156-234,1706-1714`) -- and, once `Mesh.enable_gradients_texture()` made the texture or the
vertex colours a parameter (`diffdope.py:909-920`), to that colour attribute (`ddope_render_bwd_attr`)."""
import torch


class _RenderMtx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mtx, attr, scene):
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
        d_mtx = sc.render_bwd(mtx, d_rgb, d_depth, d_mask) if ctx.needs_input_grad[0] else None
        d_attr = None
        if ctx.needs_input_grad[1] and d_rgb is not None:
            d_attr = sc.render_attr_grad(mtx, d_rgb)
        return d_mtx, d_attr, None


def render_mtx(scene, mtx, attr=None):
    """attr: the single texture [Ht,Wt,3] / vertex-colour table [V,3] the scene was built from, when it requires grad."""
    return _RenderMtx.apply(mtx, attr, scene)
