"""Registration losses of the reference, backed by the sm_100a kernels.

Mirrors util/losses.py (NCC_Loss :132-261, Grad_Loss :81-130), the N-D twins in
models/voxelmorph/torchvoxelmorph/losses.py (NCC :7-67, Grad :93-117) and the two helpers of
models/registration_model.py (smooothing_loss :25-32, calculate_L1_loss :255-263).
"""
import torch

from . import _lib


def _f32c(t):
    if t.dtype != torch.float32:
        raise _lib.DfmirError(f"dfmir_b200 kernels are fp32; got {t.dtype}")
    return t.contiguous()


_ws_cache = {}


def _workspace(nbytes, device):
    """Per-device scratch buffer (grown on demand; stream-ordered reuse on the current stream)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


class _NCCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, I, J, mask, win, eps, reduction):
        _lib.require_cuda(I, J)
        I, J = _f32c(I), _f32c(J)
        if I.shape != J.shape or I.shape[1] != 1:
            raise _lib.DfmirError(f"NCC: expected two (B,1,*S) volumes, got {tuple(I.shape)} / {tuple(J.shape)}")
        B = I.shape[0]
        shape = list(I.shape[2:])
        nd = len(shape)
        if mask is not None:
            mask = _f32c(mask.to(torch.float32).expand_as(I))
        nbytes = _lib.lib().dfmir_ncc_workspace_bytes(B, nd, _lib._ints(shape), win)
        ws = _workspace(nbytes, I.device)
        out = torch.empty(3, dtype=torch.float32, device=I.device)
        _lib.call("dfmir_ncc_fwd", I, J, mask, out, ws, _lib.size_t(ws.numel()), B, nd, shape, win, float(eps), reduction)
        ctx.save_for_backward(I, J, mask, out)
        ctx.meta = (B, nd, shape, win, float(eps), reduction)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        I, J, mask, out = ctx.saved_tensors
        B, nd, shape, win, eps, reduction = ctx.meta
        g = g.to(torch.float32).reshape(1).contiguous()
        nbytes = _lib.lib().dfmir_ncc_workspace_bytes(B, nd, _lib._ints(shape), win)
        ws = _workspace(nbytes, I.device)
        dI = dJ = None
        if ctx.needs_input_grad[0]:
            dI = torch.empty_like(I)
            _lib.call("dfmir_ncc_bwd", I, J, mask, out, g, dI, ws, _lib.size_t(ws.numel()), B, nd, shape, win, eps, reduction)
        if ctx.needs_input_grad[1]:
            dJ = torch.empty_like(J)
            _lib.call("dfmir_ncc_bwd", J, I, mask, out, g, dJ, ws, _lib.size_t(ws.numel()), B, nd, shape, win, eps, reduction)
        return dI, dJ, None, None, None, None


def _uniform_window(kernel_var, ndims):
    if kernel_var is None:
        kernel_var = [9] * ndims
    kv = [int(k) for k in kernel_var]
    if len(kv) != ndims or len(set(kv)) != 1:
        raise _lib.DfmirError(f"NCC: kernel_var {kernel_var} must be {ndims} equal odd window sizes")
    return kv[0]


class NCC_Loss(torch.nn.Module):
    """Local (windowed) normalised cross correlation, -sqrt(mean(cc)) (reference: util/losses.py:132-261)."""

    def __init__(self, device=None, kernel_var=None, name=None, kernel_type='mean', eps=1e-5, *args, **kwargs):
        super().__init__()
        self.name = 'ncc' if name is None else name
        self.device = device
        self.kernel_var = kernel_var
        self.kernel_type = kernel_type
        self.eps = eps
        assert kernel_type in ['mean', 'gaussian', 'linear']
        if kernel_type != 'mean':
            raise NotImplementedError("dfmir_b200 NCC_Loss implements the 'mean' (box) kernel the reference model uses")

    def forward(self, prediction, target, mask=None, *args, **kwargs):
        ndims = prediction.dim() - 2
        assert ndims in [1, 2, 3], "volumes should be 1 to 3 dimensions. found: %d" % ndims
        win = _uniform_window(self.kernel_var, ndims)
        return _NCCFn.apply(prediction, target, mask, win, self.eps, 0)


class NCC:
    """vxm variant: -mean(cc) (reference: models/voxelmorph/torchvoxelmorph/losses.py:7-67)."""

    def __init__(self, win=None):
        self.win = win

    def loss(self, y_true, y_pred):
        ndims = y_true.dim() - 2
        win = _uniform_window(self.win, ndims)
        # reference: Ii = y_true, Ji = y_pred
        return _NCCFn.apply(y_true, y_pred, None, win, 1e-5, 1)


class _GradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, penalty, loss_mult):
        _lib.require_cuda(x)
        x = _f32c(x)
        planes = x.shape[0] * x.shape[1]
        shape = list(x.shape[2:])
        ws = _workspace(_lib.lib().dfmir_grad_loss_workspace_bytes(), x.device)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.call("dfmir_grad_loss_fwd", x, out, ws, _lib.size_t(ws.numel()), planes, len(shape), shape, penalty, float(loss_mult))
        ctx.save_for_backward(x)
        ctx.meta = (planes, shape, penalty, float(loss_mult))
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        planes, shape, penalty, loss_mult = ctx.meta
        g = g.to(torch.float32).reshape(1).contiguous()
        dx = torch.empty_like(x)
        _lib.call("dfmir_grad_loss_bwd", x, g, dx, planes, len(shape), shape, penalty, loss_mult)
        return dx, None, None


class Grad_Loss(torch.nn.Module):
    """Finite-difference smoothness of a deformation field (reference: util/losses.py:81-130)."""

    def __init__(self, dim=2, penalty='l2', name=None, loss_mult=None, *args, **kwargs):
        super().__init__()
        self.name = 'gradient' if name is None else name
        assert dim in [2, 3]
        self.dim = dim
        self.penalty = penalty
        self.loss_mult = loss_mult

    def forward(self, prediction, *args, **kwargs):
        if 'mask' in kwargs:
            prediction = prediction * kwargs['mask']
        if prediction.dim() - 2 != self.dim:
            raise _lib.DfmirError(f"Grad_Loss(dim={self.dim}) got a {prediction.dim() - 2}-D field")
        return _GradFn.apply(prediction, 2 if self.penalty == 'l2' else 1,
                             1.0 if self.loss_mult is None else self.loss_mult)


class Grad:
    """vxm variant (reference: vxm losses.py:93-117; its indexing is 3-D only, default l1)."""

    def __init__(self, penalty='l1', loss_mult=None):
        self.penalty = penalty
        self.loss_mult = loss_mult

    def loss(self, _, y_pred):
        return _GradFn.apply(y_pred, 2 if self.penalty == 'l2' else 1, 1.0 if self.loss_mult is None else self.loss_mult)


def smooothing_loss(y_pred):
    """(mean(dx^2) + mean(dy^2)) / 2 (reference: models/registration_model.py:25-32)."""
    return _GradFn.apply(y_pred, 2, 1.0)


class _L1MaskedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, mask, mu, mv, thr):
        _lib.require_cuda(a, b)
        a, b = _f32c(a), _f32c(b)
        if a.shape != b.shape:
            raise _lib.DfmirError(f"L1: shape mismatch {tuple(a.shape)} / {tuple(b.shape)}")
        if mask is not None:
            mask = mask.expand_as(a).to(torch.uint8).contiguous()
        if mu is not None:
            mu, mv = _f32c(mu.detach()), _f32c(mv.detach())
        ws = _workspace(_lib.lib().dfmir_l1_masked_workspace_bytes(), a.device)
        out = torch.empty(2, dtype=torch.float32, device=a.device)
        _lib.call("dfmir_l1_masked_fwd", a, b, mask, mu, mv, float(thr), out, ws, _lib.size_t(ws.numel()), _lib.i64(a.numel()))
        ctx.save_for_backward(a, b, mask, mu, mv, out)
        ctx.thr = float(thr)
        msum = out[1]
        ctx.mark_non_differentiable(msum)
        return out[0], msum

    @staticmethod
    def backward(ctx, g, _gm=None):
        a, b, mask, mu, mv, out = ctx.saved_tensors
        g = g.to(torch.float32).reshape(1).contiguous()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.call("dfmir_l1_masked_bwd", a, b, mask, mu, mv, ctx.thr, out, g, da, db, _lib.i64(a.numel()))
        return da, db, None, None, None, None


def calculate_L1_loss(src, tgt, mask=None):
    """sum(|src-tgt|*mask)/sum(mask) (reference: models/registration_model.py:255-263).
    Deviation: an empty mask yields a float 0 on the device instead of the reference's host-side
    `torch.tensor(0)`, which avoids the D2H sync of `torch.sum(mask) == 0`."""
    return _L1MaskedFn.apply(src, tgt, mask, None, None, 0.0)[0]


def l1_threshold_masked(src, tgt, mu, mv, thr=-0.95, return_mask_sum=False):
    """calculate_L1_loss with the mask (mu > thr) | (mv > thr) of registration_model.py:160-161 fused in.
    return_mask_sum also returns sum(mask) as a device scalar (used for the global-batch normalisation)."""
    loss, msum = _L1MaskedFn.apply(src, tgt, None, mu, mv, thr)
    return (loss, msum) if return_mask_sum else loss
