"""`xfm_points` / `xfm_vectors`: the two ops the reference exports from its JIT-built
`renderutils_plugin` (`diffdope/ops.py:104-175`, `diffdope/c_src/*`), here calling the prebuilt
libddope_b200.so through its C ABI (`ddope_xfm_fwd/bwd/bwd_mtx/bwd_full`).

Same semantics: points [B,N,3] or [1,N,3], matrix [B,4,4]; xfm_points -> [B,N,4] = M [p,1];
xfm_vectors -> [B,N,3] = M3x3 v. `use_python=True` keeps the reference's torch.matmul validation
path (`ops.py:137-141,163-167`)."""
import ctypes

import torch

from . import _native


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _check_inputs(points, matrix):
    if not (points.is_cuda and matrix.is_cuda):
        raise RuntimeError("points and matrix must be cuda tensors")
    if points.dtype != torch.float32 or matrix.dtype != torch.float32:
        raise RuntimeError("points and matrix must be fp32")
    if points.dim() != 3 or points.shape[2] != 3:
        raise RuntimeError("points must have 3 dimensions and 3 channels")
    if matrix.dim() != 3 or matrix.shape[1] != 4 or matrix.shape[2] != 4:
        raise RuntimeError("matrix must have 3 dimensions and 4 channels")
    if points.shape[0] not in (1, matrix.shape[0]):
        raise RuntimeError("points batch must be 1 or match the matrix batch")


class _xfm_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, matrix, isPoints):
        _check_inputs(points, matrix)
        points, matrix = points.contiguous(), matrix.contiguous()
        ctx.save_for_backward(points, matrix)
        ctx.isPoints = isPoints
        B, N = matrix.shape[0], points.shape[1]
        out = torch.empty(B, N, 4 if isPoints else 3, device=points.device, dtype=torch.float32)
        _native._check(_native.lib().ddope_xfm_fwd(_p(points), points.shape[0], N, _p(matrix), B, int(isPoints), _p(out), _native._stream()))
        return out

    @staticmethod
    def backward(ctx, dout):
        points, matrix = ctx.saved_tensors
        dout = dout.contiguous()
        B, N, Bp = matrix.shape[0], points.shape[1], points.shape[0]
        L = _native.lib()
        points_grad = matrix_grad = None
        if ctx.needs_input_grad[0]:
            points_grad = torch.empty(B, N, 3, device=dout.device, dtype=torch.float32)
            _native._check(L.ddope_xfm_bwd(_p(matrix), B, N, _p(dout), int(ctx.isPoints), _p(points_grad), _native._stream()))
            if Bp == 1 and B > 1:
                points_grad = points_grad.sum(dim=0, keepdim=True)
        if ctx.needs_input_grad[1]:
            matrix_grad = torch.empty(B, 4, 4, device=dout.device, dtype=torch.float32)
            _native._check(L.ddope_xfm_bwd_mtx(_p(points), Bp, N, _p(dout), B, int(ctx.isPoints), _p(matrix_grad), _native._stream()))
        return points_grad, matrix_grad, None


def xfm_points(points, matrix, use_python=False):
    """Transform points: [B|1,N,3] x [B,4,4] -> homogeneous [B,N,4]."""
    if use_python:
        out = torch.matmul(torch.nn.functional.pad(points, pad=(0, 1), mode="constant", value=1.0), torch.transpose(matrix, 1, 2))
    else:
        out = _xfm_func.apply(points, matrix, True)
    if torch.is_anomaly_enabled():
        assert torch.all(torch.isfinite(out)), "Output of xfm_points contains inf or NaN"
    return out


def xfm_vectors(vectors, matrix, use_python=False):
    """Transform vectors: [B|1,N,3] x [B,4,4] -> [B,N,3] (rotation part only)."""
    if use_python:
        out = torch.matmul(torch.nn.functional.pad(vectors, pad=(0, 1), mode="constant", value=0.0), torch.transpose(matrix, 1, 2))[..., 0:3].contiguous()
    else:
        out = _xfm_func.apply(vectors, matrix, False)
    if torch.is_anomaly_enabled():
        assert torch.all(torch.isfinite(out)), "Output of xfm_vectors contains inf or NaN"
    return out
