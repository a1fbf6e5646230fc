"""Host-side mirror of the reference's VoxelMorph layers, backed by the sm_100a kernels.

Same constructors / forward signatures / buffers as models/voxelmorph/torchvoxelmorph/layers.py
(SpatialTransformer :6-48, VecInt :51-68, ResizeTransform :71-97) so that they drop into
VxmDense / REGISTRATIONModel unchanged.  Each forward is one C-ABI call (see include/dfmir_b200.h).
"""
import os

import torch
import torch.nn as nn

from . import _lib

# 0: reproduce the CPU ATen arithmetic (the oracle); 1: reproduce CUDA ATen (reciprocal multiply)
COORD_MODE = int(os.environ.get("DFMIR_COORD_MODE", "0"))


def _f32c(t):
    if t.dtype != torch.float32:
        raise _lib.DfmirError(f"dfmir_b200 kernels are fp32; got {t.dtype}")
    return t.contiguous()


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, flow, interp, coord_mode):
        _lib.require_cuda(src, flow)
        src, flow = _f32c(src), _f32c(flow)
        B, C = src.shape[:2]
        shape = list(flow.shape[2:])
        nd = len(shape)
        if flow.shape[0] != B or flow.shape[1] != nd or list(src.shape[2:]) != shape:
            raise _lib.DfmirError(f"warp: src {tuple(src.shape)} / flow {tuple(flow.shape)} mismatch")
        out = torch.empty_like(src)
        _lib.call("dfmir_warp_fwd", src, flow, out, None, B, C, nd, shape, interp, coord_mode)
        ctx.save_for_backward(src, flow)
        ctx.meta = (B, C, nd, shape, interp, coord_mode)
        return out

    @staticmethod
    def backward(ctx, gout):
        src, flow = ctx.saved_tensors
        B, C, nd, shape, interp, coord_mode = ctx.meta
        if interp != 0:
            raise _lib.DfmirError("warp: backward is defined for linear interpolation only")
        need_src, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_src = torch.zeros_like(src) if need_src else None
        d_flow = torch.empty_like(flow) if need_flow else None
        _lib.call("dfmir_warp_bwd", _f32c(gout), src, flow, d_src, d_flow, B, C, nd, shape, coord_mode)
        return d_src, d_flow, None, None


def warp(src, flow, mode="bilinear", coord_mode=None):
    interp = {"bilinear": 0, "nearest": 1}.get(mode)
    if interp is None:
        raise _lib.DfmirError(f"SpatialTransformer mode {mode!r} not supported (bilinear | nearest)")
    return _WarpFn.apply(src, flow, interp, COORD_MODE if coord_mode is None else coord_mode)


def warp_indices(src, flow, mode="bilinear", coord_mode=None):
    """Debug / parity entry: returns (out, idx) with idx int32 (B,nd,*S) = the integer sampling
    indices (floor for bilinear, round-half-even for nearest) the kernel used."""
    _lib.require_cuda(src, flow)
    src, flow = _f32c(src), _f32c(flow)
    B, C = src.shape[:2]
    shape = list(flow.shape[2:])
    out = torch.empty_like(src)
    idx = torch.empty(flow.shape, dtype=torch.int32, device=flow.device)
    _lib.call("dfmir_warp_fwd", src, flow, out, idx, B, C, len(shape), shape,
              {"bilinear": 0, "nearest": 1}[mode], COORD_MODE if coord_mode is None else coord_mode)
    return out, idx


class SpatialTransformer(nn.Module):
    """N-D spatial transformer (reference: layers.py:6-48). `grid` is kept as a buffer only so that
    state_dicts round-trip with the reference; the kernel derives coordinates from thread indices."""

    def __init__(self, size, mode='bilinear'):
        super().__init__()
        self.mode = mode
        vectors = [torch.arange(0, s) for s in size]
        grid = torch.stack(torch.meshgrid(*vectors, indexing='ij')).unsqueeze(0).float()
        self.register_buffer('grid', grid)

    def forward(self, src, flow):
        return warp(src, flow, self.mode)


class _VecIntFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, nsteps, bidir, coord_mode):
        _lib.require_cuda(vec)
        vec = _f32c(vec)
        B, nd = vec.shape[:2]
        shape = list(vec.shape[2:])
        if nd != len(shape):
            raise _lib.DfmirError(f"VecInt: field {tuple(vec.shape)} must have {len(shape)} channels")
        keep = bool(ctx.needs_input_grad[0])  # keep every squaring step for the backward pass
        Bv = 2 * B if bidir else B
        nslab = nsteps if keep else min(nsteps, 2)
        steps = torch.empty((nslab, Bv, nd, *shape), dtype=vec.dtype, device=vec.device)
        _lib.call("dfmir_vecint_fwd", vec, steps, B, nd, shape, nsteps, bidir, keep, coord_mode)
        out = steps[nsteps - 1 if keep else (nsteps - 1) & 1]
        if keep:
            ctx.save_for_backward(vec, steps)
        ctx.meta = (B, nd, shape, nsteps, bidir, coord_mode)
        return out

    @staticmethod
    def backward(ctx, gout):
        vec, steps = ctx.saved_tensors
        B, nd, shape, nsteps, bidir, coord_mode = ctx.meta
        gout = _f32c(gout)
        work = torch.empty((2,) + tuple(gout.shape), dtype=gout.dtype, device=gout.device)
        d_vel = torch.empty_like(vec)
        _lib.call("dfmir_vecint_bwd", gout, vec, steps, work, d_vel, B, nd, shape, nsteps, bidir, coord_mode)
        return d_vel, None, None, None


def vecint(vec, nsteps, bidir=False, coord_mode=None):
    """Scaling and squaring. bidir=True returns (2B,...) = [integrate(vec); integrate(-vec)]."""
    if nsteps == 0:
        return torch.cat([vec, -vec]) if bidir else vec
    return _VecIntFn.apply(vec, nsteps, bidir, COORD_MODE if coord_mode is None else coord_mode)


class VecInt(nn.Module):
    """Integrates a vector field via scaling and squaring (reference: layers.py:51-68)."""

    def __init__(self, inshape, nsteps):
        super().__init__()
        assert nsteps >= 0, 'nsteps should be >= 0, found: %d' % nsteps
        self.nsteps = nsteps
        self.scale = 1.0 / (2 ** self.nsteps)
        self.transformer = SpatialTransformer(inshape)

    def forward(self, vec):
        return vecint(vec, self.nsteps, bidir=False)

    def forward_bidir(self, vec):
        """Integrate +vec and -vec in one launch per step; returns (pos, neg)."""
        out = vecint(vec, self.nsteps, bidir=True)
        B = vec.shape[0]
        return out[:B], out[B:]


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_shape, pre_mul, post_mul):
        _lib.require_cuda(x)
        x = _f32c(x)
        B, C = x.shape[:2]
        in_shape = list(x.shape[2:])
        y = torch.empty((B, C, *out_shape), dtype=x.dtype, device=x.device)
        _lib.call("dfmir_resize_linear_fwd", x, y, B * C, len(in_shape), in_shape, list(out_shape),
                  float(pre_mul), float(post_mul))
        ctx.meta = (B, C, in_shape, list(out_shape), float(pre_mul), float(post_mul))
        return y

    @staticmethod
    def backward(ctx, gy):
        B, C, in_shape, out_shape, pre_mul, post_mul = ctx.meta
        gy = _f32c(gy)
        gx = torch.empty((B, C, *in_shape), dtype=gy.dtype, device=gy.device)
        _lib.call("dfmir_resize_linear_bwd", gy, gx, B * C, len(in_shape), in_shape, out_shape, pre_mul, post_mul)
        return gx, None, None, None


class ResizeTransform(nn.Module):
    """Resize a transform: resample the vector field and rescale it (reference: layers.py:71-97)."""

    def __init__(self, vel_resize, ndims):
        super().__init__()
        self.factor = 1.0 / vel_resize
        self.mode = 'linear'
        if ndims == 2:
            self.mode = 'bi' + self.mode
        elif ndims == 3:
            self.mode = 'tri' + self.mode

    def forward(self, x):
        if self.factor == 1:
            return x
        # F.interpolate(scale_factor=f) output size: floor(in * f)
        out_shape = [int(s * self.factor) for s in x.shape[2:]]
        if self.factor < 1:   # resize first, then rescale (layers.py:86-89)
            return _ResizeFn.apply(x, out_shape, 1.0, self.factor)
        return _ResizeFn.apply(x, out_shape, self.factor, 1.0)  # rescale first (layers.py:91-94)
