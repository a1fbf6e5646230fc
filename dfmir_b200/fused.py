"""Fused registration tail: integrate -> resize -> warp -> NCC + Grad in ONE cooperative launch
(csrc/fused_reg.cu), the B200 counterpart of the ~130 kernels the reference runs for
VxmDense.forward's tail (models/voxelmorph/torchvoxelmorph/networks.py:1129-1139) plus
NCC_Loss / Grad_Loss (util/losses.py:81-261).

    warped, flow, ncc, grad = integrate_warp_loss(vel, moving, fixed, nsteps=7, win=9)

`vel` is the half-resolution stationary velocity field (B, nd, *S/2) (the output of VxmDense's flow head after
ResizeTransform(int_downsize)); `flow` is the integrated full-resolution displacement (what the reference's
VxmDense returns as pos_flow), `warped` the moved image.  All four outputs are differentiable: the backward
pass chains the stand-alone kernels (ncc_bwd, grad_loss_bwd, warp_bwd, resize_bwd, vecint_bwd).
"""
import torch

from . import _lib
from .layers import COORD_MODE, _f32c
from .losses import _workspace


class _FusedRegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vel, moving, fixed, nsteps, win, eps, ncc_reduction, grad_penalty, grad_mult, coord_mode):
        _lib.require_cuda(vel, moving, fixed)
        vel, moving, fixed = _f32c(vel), _f32c(moving), _f32c(fixed)
        B, nd = vel.shape[:2]
        half = list(vel.shape[2:])
        full = [2 * s for s in half]
        C = moving.shape[1]
        if nd != len(half) or list(moving.shape[2:]) != full or tuple(fixed.shape) != (B, 1, *full) or C != 1:
            raise _lib.DfmirError(f"integrate_warp_loss: vel {tuple(vel.shape)}, moving {tuple(moving.shape)}, fixed "
                                  f"{tuple(fixed.shape)}: need (B,nd,*S/2), (B,1,*S), (B,1,*S)")
        steps = torch.empty((nsteps, B, nd, *half), dtype=vel.dtype, device=vel.device)
        flow = torch.empty((B, nd, *full), dtype=vel.dtype, device=vel.device)
        warped = torch.empty_like(moving)
        out = torch.empty(4, dtype=torch.float32, device=vel.device)
        ws = _workspace(_lib.lib().dfmir_fused_reg_workspace_bytes(), vel.device)
        _lib.call("dfmir_fused_reg_fwd", vel, moving, fixed, steps, flow, warped, out, ws, _lib.size_t(ws.numel()),
                  B, C, nd, half, nsteps, win, float(eps), ncc_reduction, grad_penalty, float(grad_mult), coord_mode)
        ctx.save_for_backward(vel, moving, fixed, steps, flow, warped, out)
        ctx.meta = (B, C, nd, half, full, nsteps, win, float(eps), ncc_reduction, grad_penalty, float(grad_mult), coord_mode)
        return warped, flow, out[0], out[3]

    @staticmethod
    def backward(ctx, g_warped, g_flow, g_ncc, g_grad):
        vel, moving, fixed, steps, flow, warped, out = ctx.saved_tensors
        B, C, nd, half, full, nsteps, win, eps, ncc_reduction, grad_penalty, grad_mult, coord_mode = ctx.meta
        dev = vel.device
        # d loss / d warped
        d_warped = torch.zeros_like(warped) if g_warped is None else _f32c(g_warped).clone()
        if g_ncc is not None:
            nbytes = _lib.lib().dfmir_ncc_workspace_bytes(B, nd, _lib._ints(full), win)
            ws = _workspace(nbytes, dev)
            dI = torch.empty_like(warped)
            _lib.call("dfmir_ncc_bwd", warped, fixed, None, out[:3].contiguous(), g_ncc.to(torch.float32).reshape(1).contiguous(),
                      dI, ws, _lib.size_t(ws.numel()), B, nd, full, win, eps, ncc_reduction)
            d_warped += dI
        # d loss / d flow (full resolution)
        d_flow = torch.zeros_like(flow) if g_flow is None else _f32c(g_flow).clone()
        if g_grad is not None:
            dg = torch.empty_like(flow)
            _lib.call("dfmir_grad_loss_bwd", flow, g_grad.to(torch.float32).reshape(1).contiguous(), dg, B * nd, nd, full,
                      grad_penalty, grad_mult)
            d_flow += dg
        need_moving = ctx.needs_input_grad[1]
        d_moving = torch.zeros_like(moving) if need_moving else None
        dw = torch.empty_like(flow)
        _lib.call("dfmir_warp_bwd", d_warped, moving, flow, d_moving, dw, B, C, nd, full, coord_mode)
        d_flow += dw
        d_vel = None
        if ctx.needs_input_grad[0]:
            d_field = torch.empty((B, nd, *half), dtype=vel.dtype, device=dev)
            _lib.call("dfmir_resize_linear_bwd", d_flow, d_field, B * nd, nd, half, full, 2.0, 1.0)
            work = torch.empty((2, B, nd, *half), dtype=vel.dtype, device=dev)
            d_vel = torch.empty_like(vel)
            _lib.call("dfmir_vecint_bwd", d_field, vel, steps, work, d_vel, B, nd, half, nsteps, 0, coord_mode)
        d_fixed = None
        if ctx.needs_input_grad[2] and g_ncc is not None:
            nbytes = _lib.lib().dfmir_ncc_workspace_bytes(B, nd, _lib._ints(full), win)
            ws = _workspace(nbytes, dev)
            d_fixed = torch.empty_like(fixed)
            _lib.call("dfmir_ncc_bwd", fixed, warped, None, out[:3].contiguous(), g_ncc.to(torch.float32).reshape(1).contiguous(),
                      d_fixed, ws, _lib.size_t(ws.numel()), B, nd, full, win, eps, ncc_reduction)
        return d_vel, d_moving, d_fixed, None, None, None, None, None, None, None


def integrate_warp_loss(vel, moving, fixed, nsteps=7, win=9, eps=1e-5, ncc="sqrt_mean", grad_penalty="l2",
                        grad_mult=1.0, coord_mode=None):
    """One launch: flow = fullsize(VecInt(vel)); warped = SpatialTransformer(moving, flow);
    ncc = NCC_Loss(kernel_var=[win]*nd)(warped, fixed) (`ncc="mean"`: the vxm NCC, -mean cc);
    grad = Grad_Loss(dim=nd, penalty)(flow) * grad_mult.  Returns (warped, flow, ncc, grad)."""
    red = {"sqrt_mean": 0, "mean": 1}[ncc]
    pen = {"l1": 1, "l2": 2}[grad_penalty]
    return _FusedRegFn.apply(vel, moving, fixed, int(nsteps), int(win), float(eps), red, pen, float(grad_mult),
                             COORD_MODE if coord_mode is None else coord_mode)
