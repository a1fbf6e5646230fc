"""autograd bindings of the convolution / normalisation / resampling / PatchNCE entry points of
libdfmir_b200.so (include/dfmir_b200.h).  torch supplies device memory, streams and the autograd
tape; every forward and backward below is a C-ABI call into hand-written sm_100a kernels.

Internal activation layout is channels-last: (N, *spatial, C) contiguous fp32.
"""
import ctypes
import weakref
import math
import os

import torch

from . import _lib

ACT_NONE, ACT_LEAKY, ACT_TANH, ACT_RELU = 0, 1, 2, 3

# convolution engine: "simt" = fp32 CUDA cores (exact), "umma" = tcgen05 tensor cores (TF32 operands)
# for the layer shapes it supports, "auto" = umma where supported.
CONV_ENGINE = os.environ.get("DFMIR_CONV_ENGINE", "auto")
# layers with fewer output positions than this stay on the fp32 kernels (launch-bound; TMA descriptor set-up
# would dominate).  Tests set it to 0 to drive small shapes through the tensor-core kernels.
UMMA_MIN_POSITIONS = int(os.environ.get("DFMIR_UMMA_MIN_POSITIONS", "4096"))
# ... unless their reduction is long: the deep U-Net levels (128 -> 64 channels, 3x3x3, at 8^3 / 4^3 voxels) have a few
# hundred positions but K = 3456; the fp32 kernel runs them on 1 - 8 CTAs for ~0.3 ms each, the tensor-core kernel in a few us
UMMA_MIN_K_SMALL = int(os.environ.get("DFMIR_UMMA_MIN_K_SMALL", "512"))
UMMA_MIN_POSITIONS_SMALL = 16


def _tc_size(positions, k_len):
    """Layer big enough for the tensor-core kernels?  positions = output positions, k_len = reduction length."""
    return positions >= UMMA_MIN_POSITIONS or (UMMA_MIN_POSITIONS > 0 and positions >= UMMA_MIN_POSITIONS_SMALL and k_len >= UMMA_MIN_K_SMALL)


class ConvProfile:
    """bench.py instrumentation: a CUDA-event pair around every convolution kernel launch (the C-ABI
    call only, on the launching stream), keyed by kernel kind, with the algorithmic FLOPs (2*M*N*K)
    of each call.  Kinds: umma_fwd / umma_dgrad / umma_wgrad (tcgen05 engine), simt (fp32 engine,
    including the direct 7x7 stem / head kernels)."""

    def __init__(self):
        self.events = []          # (kind, start, end, flops, operand + result bytes)
        self.calls, self.umma_calls = 0, 0

    def run(self, fn, flops, kind, nbytes=0.0):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        self.events.append((kind, s, e, flops, nbytes))
        self.calls += 1
        self.umma_calls += int(kind.startswith("umma"))

    def by_kind(self):
        """{kind: (total ms, total flops, calls, total algorithmic bytes)}"""
        torch.cuda.synchronize()
        out = {}
        for kind, s, e, fl, nb in self.events:
            ms, f, n, b = out.get(kind, (0.0, 0.0, 0, 0.0))
            out[kind] = (ms + s.elapsed_time(e), f + fl, n + 1, b + nb)
        return out

    def total(self):
        k = self.by_kind()
        return sum(v[0] for v in k.values()), sum(v[1] for v in k.values()), self.calls


PROFILE = None


def _run(fn, flops, kind="simt", nbytes=0.0):
    if PROFILE is None:
        fn()
    else:
        PROFILE.run(fn, flops, kind, nbytes)


def _nbytes(*tensors):
    """fp32 bytes of the operands and the result of one launch (the algorithmic traffic of a convolution)"""
    return 4.0 * sum(t.numel() for t in tensors if t is not None)


class ConvDesc(ctypes.Structure):
    _fields_ = [("nd", ctypes.c_int), ("N", ctypes.c_int), ("Cin", ctypes.c_int), ("Cout", ctypes.c_int),
                ("in_shape", ctypes.c_int * 3), ("out_shape", ctypes.c_int * 3), ("kernel", ctypes.c_int * 3),
                ("pad", ctypes.c_int * 3), ("stride", ctypes.c_int), ("act", ctypes.c_int),
                ("x_strides", ctypes.c_longlong * 5), ("y_strides", ctypes.c_longlong * 5)]


def _f32(t):
    if t.dtype != torch.float32:
        raise _lib.DfmirError(f"dfmir_b200 kernels are fp32; got {t.dtype}")
    return t


def _cl_strides(t, nd, planar=False):
    """element strides {n, spatial[nd], c} of a channels-last (N,*S,C) or planar (N,C,*S) tensor"""
    st = t.stride()
    if planar:
        return [st[0]] + list(st[2:2 + nd]) + [st[1]]
    return list(st[:nd + 2])


def _make_desc(nd, N, Cin, Cout, in_shape, out_shape, kernel, pad, stride, act, xs, ys):
    d = ConvDesc()
    d.nd, d.N, d.Cin, d.Cout, d.stride, d.act = nd, N, Cin, Cout, stride, act
    for i in range(nd):
        d.in_shape[i], d.out_shape[i], d.kernel[i], d.pad[i] = in_shape[i], out_shape[i], kernel[i], pad[i]
    for i in range(nd + 2):
        d.x_strides[i], d.y_strides[i] = xs[i], ys[i]
    return d


_ws_cache = {}


def workspace(nbytes, device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _use_umma(d, x, y_planar, positions):
    """Forward on the tensor cores?  (the backward products are decided one by one in _ConvFn.backward)"""
    taps = 1
    for i in range(d.nd):
        taps *= d.kernel[i]
    if CONV_ENGINE == "simt" or not _tc_size(positions, d.Cin * taps) or x.data_ptr() % 16 or d.Cin == 1 or d.Cout == 1:
        return False      # tiny layers are launch-bound; the 7x7 stem / head have their own exact direct kernels
    from . import umma
    return umma.supported(d, False)


class _ConvFn(torch.autograd.Function):
    """y = act(conv(x, w) + b); x (N,*S,Cin) channels-last (any strides), w (taps,Cin,Cout)."""

    @staticmethod
    def forward(ctx, x, w, bias, kernel, stride, pad, act, planar_out, bias_slot=None, res_slot=None, out_shape=None,
                flop_scale=1.0, stats_slot=None):
        ctx.bias_slot = bias_slot if act == ACT_NONE else None
        ctx.res_slot = res_slot
        _lib.require_cuda(x, w)
        x, w = _f32(x), _f32(w).contiguous()
        nd = len(kernel)
        N, Cin = x.shape[0], x.shape[-1]
        S = list(x.shape[1:1 + nd])
        Cout = w.shape[2]
        if w.shape[0] != math.prod(kernel) or w.shape[1] != Cin:
            raise _lib.DfmirError(f"conv: weight {tuple(w.shape)} does not match kernel {kernel} / Cin {Cin}")
        # out_shape: `pad` is the LEADING padding only (space-to-depth form of a strided layer, conv_cl)
        O = list(out_shape) if out_shape is not None else [(S[i] + 2 * pad[i] - kernel[i]) // stride + 1 for i in range(nd)]
        if planar_out:
            y = torch.empty((N, Cout, *O), dtype=x.dtype, device=x.device)
        else:
            y = torch.empty((N, *O, Cout), dtype=x.dtype, device=x.device)
        d = _make_desc(nd, N, Cin, Cout, S, O, kernel, pad, stride, act, _cl_strides(x, nd), _cl_strides(y, nd, planar_out))
        if bias is not None:
            bias = _f32(bias).contiguous()
        engine = "simt"
        flops = 2.0 * N * math.prod(O) * Cout * w.shape[0] * Cin * flop_scale      # algorithmic (structural zeros not counted)
        if _use_umma(d, x, planar_out, N * math.prod(O)):
            from . import umma
            rows = 0
            if stats_slot is not None and STATS_IN_EPILOGUE and act == ACT_NONE and not planar_out:
                rows = int(_lib.lib().dfmir_conv_umma_stat_rows(ctypes.byref(d)))
            if rows > 0:
                stats_slot.rows = torch.empty((N, rows, Cout, 2), dtype=x.dtype, device=x.device)
                stats_slot.rows_per_image = rows
                umma.conv_fwd(x, w, bias, y, d, flops, stat_rows=stats_slot.rows)
            else:
                umma.conv_fwd(x, w, bias, y, d, flops)
            engine = "umma"
        else:
            _run(lambda: _lib.call("dfmir_conv_fwd", x, w, bias, y, ctypes.byref(d)), flops)
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        ctx.meta = (nd, N, Cin, Cout, S, O, list(kernel), list(pad), stride, act, planar_out, bias is not None, engine, flops)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        nd, N, Cin, Cout, S, O, kernel, pad, stride, act, planar_out, has_bias, engine, flops = ctx.meta
        dy = _f32(dy)
        if act != ACT_NONE:
            dy = dy.contiguous()
            g = torch.empty_like(y)
            _lib.call("dfmir_act_bwd", y, dy, g, _lib.i64(y.numel()), act)
            dy = g
        elif CONV_ENGINE != "simt" and not planar_out:
            dy = dy.contiguous()
        dx = dw = db = None
        ys = _cl_strides(dy, nd, planar_out)
        # tensor-core engine, decided product by product (not for the stem / head: direct fp32 kernels)
        tc = CONV_ENGINE != "simt" and _tc_size(N * math.prod(O), Cin * math.prod(kernel)) and Cin != 1 and Cout != 1
        if tc:
            from . import umma
        if tc and planar_out and Cout % 4 and stride == 1 and Cin % 4 == 0 and x.data_ptr() % 16 == 0:
            # planar flow head (2 or 3 output channels): its gradient as a channels-last copy padded to 4 channels,
            # so that both backward products run on the tensor-core kernels (TMA rows must be 16-byte multiples)
            return _ConvFn._backward_padded_head(ctx, x, w, dy) + (None, None)
        pending = None
        if ctx.res_slot is not None and ctx.res_slot.dres is not None:
            pending, ctx.res_slot.dres = ctx.res_slot.dres, None         # gradient of the skip connection, shaped like x
        if ctx.needs_input_grad[0]:
            acc = (pending is not None and tc and tuple(pending.shape) == (N, *S, Cin) and pending.is_contiguous()
                   and dy.data_ptr() % 16 == 0)
            dx = pending if acc else torch.empty((N, *S, Cin), dtype=x.dtype, device=x.device)
            d = _make_desc(nd, N, Cin, Cout, S, O, kernel, pad, stride, ACT_NONE, _cl_strides(dx, nd), ys)
            if acc and _lib.lib().dfmir_conv_umma_dgrad_acc_supported(ctypes.byref(d)):
                _run(lambda: _lib.call("dfmir_conv_umma_dgrad_acc", dy, w, dx, ctypes.byref(d)), flops, "umma_dgrad",
                     _nbytes(dy, w, dx, dx))
                pending = None
            else:
                if acc:
                    dx = torch.empty((N, *S, Cin), dtype=x.dtype, device=x.device)
                if tc and dy.data_ptr() % 16 == 0 and umma.supported(d, True):
                    umma.conv_dgrad(dy, w, dx, d, flops)
                else:
                    wt = w.transpose(1, 2).contiguous()
                    _run(lambda: _lib.call("dfmir_conv_dgrad", dy, wt, dx, ctypes.byref(d)), flops)
            if pending is not None:
                dx = dx + pending
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dw = torch.zeros_like(w)
            db = torch.zeros(Cout, dtype=w.dtype, device=w.device) if has_bias else None
            slot, db_given = ctx.bias_slot, None
            if slot is not None and slot.db is not None:        # summed by the InstanceNorm backward that produced dy
                db_given, slot.db, db = slot.db, None, None
            d = _make_desc(nd, N, Cin, Cout, S, O, kernel, pad, stride, ACT_NONE, _cl_strides(x, nd), ys)
            if tc and x.data_ptr() % 16 == 0 and dy.is_contiguous():
                umma.conv_wgrad(x, dy, dw, db, d, flops)      # tcgen05 where the shape fits, else the fp32 kernel
            else:
                _run(lambda: _lib.call("dfmir_conv_wgrad", x, dy, dw, db, ctypes.byref(d)), flops)
            if db_given is not None:
                db = db_given if has_bias else None
        return dx, dw, db, None, None, None, None, None, None, None, None, None, None


def _backward_padded_head(ctx, x, w, dy):
    from . import umma
    nd, N, Cin, Cout, S, O, kernel, pad, stride, act, planar_out, has_bias, engine, flops = ctx.meta
    Cp = (Cout + 3) // 4 * 4
    dyp = torch.zeros((N, *O, Cp), dtype=dy.dtype, device=dy.device)
    dyp[..., :Cout] = dy.movedim(1, -1)
    ysp = _cl_strides(dyp, nd)
    dx = dw = db = None
    if ctx.needs_input_grad[0]:
        dx = torch.empty((N, *S, Cin), dtype=x.dtype, device=x.device)
        d = _make_desc(nd, N, Cin, Cp, S, O, kernel, pad, stride, ACT_NONE, _cl_strides(dx, nd), ysp)
        wp = torch.nn.functional.pad(w, (0, Cp - Cout))          # (taps, Cin, Cp): K-major rows of 16 bytes
        if umma.supported(d, True):
            umma.conv_dgrad(dyp, wp, dx, d, flops)
        else:
            _run(lambda: _lib.call("dfmir_conv_dgrad", dyp, wp.transpose(1, 2).contiguous(), dx, ctypes.byref(d)), flops)
    if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
        dwp = torch.zeros((w.shape[0], Cin, Cp), dtype=w.dtype, device=w.device)
        d = _make_desc(nd, N, Cin, Cp, S, O, kernel, pad, stride, ACT_NONE, _cl_strides(x, nd), ysp)
        umma.conv_wgrad(x, dyp, dwp, None, d, flops)
        dw = dwp[..., :Cout].contiguous()
        db = dy.sum(dim=[0] + list(range(2, nd + 2))) if has_bias else None
    return dx, dw, db, None, None, None, None, None, None, None, None


_ConvFn._backward_padded_head = staticmethod(_backward_padded_head)


class _PackFn(torch.autograd.Function):
    """(Cout, Cin, *k) parameter -> kernel layout (taps, Cin + pad, Cout); pad = zero rows for zero-padded activation
    channels (upsample_concat_cl(pad_channels_to=4))."""

    @staticmethod
    def forward(ctx, weight, pad_cin):
        Cout, Cin = weight.shape[:2]
        w = weight.reshape(Cout, Cin, -1).permute(2, 1, 0)
        if pad_cin > 0:
            w = torch.nn.functional.pad(w, (0, 0, 0, pad_cin))
        ctx.meta = (tuple(weight.shape), Cin, id(weight))
        return w.contiguous()

    @staticmethod
    def backward(ctx, dw):
        shape, Cin, key = ctx.meta
        _pack_cache.pop(key, None)          # the tape that used this copy is done: re-pack on the next forward
        return dw[:, :Cin, :].permute(2, 1, 0).reshape(shape), None


_pack_cache = {}


def packed_weight(weight, pad_cin=0):
    """Kernel-layout copy of a convolution weight, shared by every use of the same (unmodified) parameter inside one
    autograd tape: the generator's weights are used by four passes per step, which then cost one re-layout and one
    gradient accumulation instead of four."""
    grad = torch.is_grad_enabled() and weight.requires_grad
    key = id(weight)
    hit = _pack_cache.get(key)
    if hit is not None:
        ref, version, pad, g, w = hit
        if ref() is weight and version == weight._version and pad == pad_cin and g == grad:
            return w
    if len(_pack_cache) > 256:              # views (linear()) and replaced parameters leave dead entries behind
        for k in [k for k, v in _pack_cache.items() if v[0]() is None]:
            del _pack_cache[k]
    w = _PackFn.apply(weight, pad_cin)
    _pack_cache[key] = (weakref.ref(weight), weight._version, pad_cin, grad, w)
    return w


class BiasGradSlot:
    """Hand-over of a convolution's bias gradient from the InstanceNorm that consumes its output: the IN backward
    writes dx, which IS that convolution's output gradient, and sums it per channel on the way
    (dfmir_instnorm_bwd_bias); the convolution's backward then skips its own pass over dx.  Pass the same slot to
    conv_cl(..., bias_slot=) and instnorm_cl(..., bias_slot=); only valid when the IN is the ONLY consumer of the
    convolution's output (no feature tap on it)."""
    __slots__ = ("db",)

    def __init__(self):
        self.db = None


class StatsSlot:
    """Hand-over of InstanceNorm statistics from the convolution in front of it: the tensor-core forward kernel sums its
    result per channel while 32-voxel blocks are in registers (dfmir_conv_umma_fwd_stats) and parks the row sums here;
    instnorm_cl(..., stats_slot=) reduces them (dfmir_instnorm_fwd_rows) instead of reading the whole tensor again.
    Forward-only: the norm's backward is unchanged."""
    __slots__ = ("rows", "rows_per_image")

    def __init__(self):
        self.rows, self.rows_per_image = None, 0


STATS_IN_EPILOGUE = os.environ.get("DFMIR_IN_STATS_EPILOGUE", "1") != "0"


class ResidualGradSlot:
    """Hand-over of the residual branch's gradient inside a ResnetBlock (out = x + conv_block(x)): the last
    InstanceNorm's backward writes d(out)/d(x) of the skip connection into a buffer shaped like the block input and
    parks it here instead of returning it; the first convolution's data gradient is then ADDED to that buffer by the
    kernel's epilogue (dfmir_conv_umma_dgrad_acc) or, on the other engines, by one add.  Pass the same slot to the
    block's first conv_cl(..., res_slot=) and to instnorm_cl(..., res=x, res_slot=)."""
    __slots__ = ("dres",)

    def __init__(self):
        self.dres = None


class _S2DFn(torch.autograd.Function):
    """x (N,*S,C) -> (N,*S/2, 2^nd * C): channel = parity(d,h,w) * C + c  (dfmir_space_to_depth)."""

    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        x = _f32(x).contiguous()
        nd = x.dim() - 2
        O = [s // 2 for s in x.shape[1:1 + nd]]
        C = x.shape[-1]
        y = torch.empty((x.shape[0], *O, (1 << nd) * C), dtype=x.dtype, device=x.device)
        _lib.call("dfmir_space_to_depth", x, y, x.shape[0], nd, O, C, 0)
        ctx.meta = (tuple(x.shape), nd, O, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, nd, O, C = ctx.meta
        dy = _f32(dy).contiguous()
        dx = torch.empty(shape, dtype=dy.dtype, device=dy.device)
        _lib.call("dfmir_space_to_depth", dy, dx, shape[0], nd, O, C, 1)
        return dx


def space_to_depth_cl(x):
    return _S2DFn.apply(x)


class _PackS2DFn(torch.autograd.Function):
    """(Cout, Cin, 3^nd) parameter of a stride-2, pad-1 convolution -> kernel layout (2^nd taps, 2^nd * Cin, Cout) of the
    equivalent stride-1 convolution over the space-to-depth activation: tap k of an axis sits at (t, p) with
    k + 1 = 2 t + p, the slot (t, p) = (0, 0) is a structural zero."""

    @staticmethod
    def forward(ctx, weight):
        Cout, Cin = weight.shape[:2]
        nd = weight.dim() - 2
        w4 = torch.nn.functional.pad(weight, (1, 0) * nd)                     # (Cout, Cin, 4, ...): index k + 1
        w4 = w4.reshape(Cout, Cin, *([2, 2] * nd))                            # (.., t_a, p_a, ..)
        t_axes = [2 + 2 * a for a in range(nd)]
        p_axes = [3 + 2 * a for a in range(nd)]
        w = w4.permute(*t_axes, *p_axes, 1, 0).reshape(1 << nd, (1 << nd) * Cin, Cout)
        ctx.meta = (tuple(weight.shape), nd, id(weight))
        return w.contiguous()

    @staticmethod
    def backward(ctx, dw):
        shape, nd, key = ctx.meta
        _pack_cache.pop((key, "s2d"), None)
        Cout, Cin = shape[:2]
        g = dw.reshape(*([2] * nd), *([2] * nd), Cin, Cout)                    # (t.., p.., Cin, Cout)
        order = [2 * nd + 1, 2 * nd]
        for a in range(nd):
            order += [a, nd + a]
        g = g.permute(*order).reshape(Cout, Cin, *([4] * nd))
        return g[(slice(None), slice(None)) + (slice(1, None),) * nd]


def packed_weight_s2d(weight):
    grad = torch.is_grad_enabled() and weight.requires_grad
    key = (id(weight), "s2d")
    hit = _pack_cache.get(key)
    if hit is not None:
        ref, version, pad, g, w = hit
        if ref() is weight and version == weight._version and g == grad:
            return w
    w = _PackS2DFn.apply(weight)
    _pack_cache[key] = (weakref.ref(weight), weight._version, 0, grad, w)
    return w


# stride-2 3^nd encoder layers as stride-1 2^nd convolutions over the space-to-depth activation (tensor-core engine)
S2D_STRIDED = os.environ.get("DFMIR_S2D_STRIDED", "1") != "0"


def conv_cl(x, weight, bias, stride=1, pad=0, act=ACT_NONE, planar_out=False, bias_slot=None, res_slot=None, stats_slot=None):
    """Convolution on a channels-last activation with a PyTorch-layout weight (Cout, Cin, *k).
    Returns (N,*O,Cout), or the planar (N,Cout,*O) when planar_out."""
    nd = weight.dim() - 2
    kernel = list(weight.shape[2:])
    pads = [pad] * nd if isinstance(pad, int) else list(pad)
    Cin = weight.shape[1]
    S = list(x.shape[1:1 + nd])
    if (stride == 2 and S2D_STRIDED and CONV_ENGINE != "simt" and kernel == [3] * nd and pads == [1] * nd
            and all(s % 2 == 0 for s in S) and x.shape[-1] == Cin and weight.shape[0] % 4 == 0 and weight.shape[0] >= 16
            and _tc_size(x.shape[0] * math.prod(S) // (1 << nd), Cin * 3 ** nd) and not planar_out):
        xs = space_to_depth_cl(x)
        return _ConvFn.apply(xs, packed_weight_s2d(weight), bias, [2] * nd, 1, [1] * nd, act, False, bias_slot, res_slot,
                             [s // 2 for s in S], 27.0 / 64.0 if nd == 3 else 9.0 / 16.0)
    w = packed_weight(weight, x.shape[-1] - Cin)
    return _ConvFn.apply(x, w, bias, kernel, stride, pads, act, planar_out, bias_slot, res_slot, None, 1.0, stats_slot)


class _InstNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, relu, out_pad, res_pad, eps, bias_slot=None, res_slot=None, stats_slot=None):
        ctx.bias_slot = bias_slot
        ctx.res_slot = res_slot if res is not None else None
        _lib.require_cuda(x)
        x = _f32(x).contiguous()
        N, H, W, C = x.shape
        if res is not None:
            res = _f32(res).contiguous()
            if tuple(res.shape) != (N, H + 2 * res_pad, W + 2 * res_pad, C):
                raise _lib.DfmirError(f"instnorm: residual {tuple(res.shape)} does not match {tuple(x.shape)} pad {res_pad}")
        y = torch.empty((N, H + 2 * out_pad, W + 2 * out_pad, C), dtype=x.dtype, device=x.device)
        stats = torch.empty((N, C, 2), dtype=x.dtype, device=x.device)
        rows = None
        if stats_slot is not None and stats_slot.rows is not None:
            rows, stats_slot.rows = stats_slot.rows, None
            if tuple(rows.shape) != (N, stats_slot.rows_per_image, C, 2):
                raise _lib.DfmirError(f"instnorm: statistic rows {tuple(rows.shape)} do not belong to {tuple(x.shape)}")
        if rows is not None:
            _lib.call("dfmir_instnorm_fwd_rows", x, res, y, stats, rows, rows.shape[1], N, H, W, C, float(eps), int(relu),
                      out_pad, res_pad)
        else:
            nbytes = _lib.lib().dfmir_instnorm_workspace_bytes(N, C)
            ws = workspace(nbytes, x.device)
            _lib.call("dfmir_instnorm_fwd", x, res, y, stats, ws, _lib.size_t(ws.numel()), N, H, W, C, float(eps),
                      int(relu), out_pad, res_pad)
        ctx.save_for_backward(x, stats)
        ctx.meta = (N, H, W, C, int(relu), out_pad, res_pad, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        N, H, W, C, relu, out_pad, res_pad, has_res = ctx.meta
        if dy.is_sparse:          # the layer's only consumer was a PatchNCE tap (last layer of an encoder-only pass)
            dy = dy.to_dense()
        dy = _f32(dy).contiguous()
        dx = torch.empty_like(x)
        dres = None
        if has_res and ctx.needs_input_grad[1]:
            dres = torch.empty((N, H + 2 * res_pad, W + 2 * res_pad, C), dtype=x.dtype, device=x.device)
        ws = workspace(_lib.lib().dfmir_instnorm_workspace_bytes(N, C), x.device)
        slot = ctx.bias_slot
        dbias = None
        # the bias by-product exists on the 128-bit kernels only (csrc/norm_resample.cu in_v4_ok)
        v4 = (C % 4 == 0 and C // 4 <= 256 and 256 % (C // 4) == 0 and N <= 65535
              and all(t is None or t.data_ptr() % 16 == 0 for t in (dy, x, dx, dres, stats)))
        if slot is not None and v4:
            dbias = torch.empty(C, dtype=x.dtype, device=x.device)
        _lib.call("dfmir_instnorm_bwd_bias", dy, x, stats, dx, dres, dbias, ws, _lib.size_t(ws.numel()), N, H, W, C, relu,
                  out_pad, res_pad)
        if slot is not None:
            slot.db = dbias
        if ctx.res_slot is not None and dres is not None:
            ctx.res_slot.dres, dres = dres, None        # taken up by the block's first convolution backward
        return dx, dres, None, None, None, None, None, None, None


def instnorm_cl(x, relu=False, out_pad=0, res=None, res_pad=0, eps=1e-5, bias_slot=None, res_slot=None, stats_slot=None):
    """InstanceNorm2d(affine=False) [+ReLU] [+res] written with a reflected halo of width out_pad."""
    return _InstNormFn.apply(x, res, relu, out_pad, res_pad, eps, bias_slot, res_slot, stats_slot)


class _PadReflectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad):
        _lib.require_cuda(x)
        x = _f32(x).contiguous()
        N, H, W, C = x.shape
        y = torch.empty((N, H + 2 * pad, W + 2 * pad, C), dtype=x.dtype, device=x.device)
        _lib.call("dfmir_pad_reflect_fwd", x, y, N, H, W, C, pad)
        ctx.meta = (N, H, W, C, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C, pad = ctx.meta
        dy = _f32(dy).contiguous()
        dx = torch.empty((N, H, W, C), dtype=dy.dtype, device=dy.device)
        _lib.call("dfmir_pad_reflect_bwd", dy, dx, N, H, W, C, pad)
        return dx, None


def pad_reflect_cl(x, pad):
    return _PadReflectFn.apply(x, pad)


class _BlurFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, up):
        _lib.require_cuda(x)
        x = _f32(x).contiguous()
        N, H, W, C = x.shape
        if up:
            y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
            _lib.call("dfmir_blur_up_fwd", x, y, N, H, W, C)
        else:
            y = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=x.dtype, device=x.device)
            _lib.call("dfmir_blur_down_fwd", x, y, N, H, W, C)
        ctx.meta = (N, H, W, C, up)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C, up = ctx.meta
        dy = _f32(dy).contiguous()
        dx = torch.empty((N, H, W, C), dtype=dy.dtype, device=dy.device)
        _lib.call("dfmir_blur_up_bwd" if up else "dfmir_blur_down_bwd", dy, dx, N, H, W, C)
        return dx, None


def blur_down_cl(x):
    return _BlurFn.apply(x, False)


def blur_up_cl(x):
    return _BlurFn.apply(x, True)


class _UpCatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, Cs):
        _lib.require_cuda(a, b)
        a, b = _f32(a).contiguous(), _f32(b).contiguous()
        nd = b.dim() - 2
        N, shape, C1, C2 = b.shape[0], list(b.shape[1:1 + nd]), a.shape[-1], b.shape[-1]
        if [2 * s for s in a.shape[1:1 + nd]] != shape:
            raise _lib.DfmirError(f"upsample_concat: {tuple(a.shape)} x2 does not match skip {tuple(b.shape)}")
        y = torch.empty((N, *shape, Cs), dtype=a.dtype, device=a.device)
        _lib.call("dfmir_upsample_concat_padded_fwd", a, b, y, N, nd, shape, C1, C2, Cs)
        ctx.meta = (N, nd, shape, C1, C2, Cs, tuple(a.shape), tuple(b.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        N, nd, shape, C1, C2, Cs, sa, sb = ctx.meta
        dy = _f32(dy).contiguous()
        da = torch.empty(sa, dtype=dy.dtype, device=dy.device) if ctx.needs_input_grad[0] else None
        db = torch.empty(sb, dtype=dy.dtype, device=dy.device) if ctx.needs_input_grad[1] else None
        if da is not None or db is not None:
            _lib.call("dfmir_upsample_concat_padded_bwd", dy, da, db, N, nd, shape, C1, C2, Cs)
        return da, db, None


def upsample_concat_cl(a, b, pad_channels_to=1):
    """cat([nearest_x2(a), b], channel) for channels-last tensors (U-Net skip connection).  With
    pad_channels_to = 4 the result carries zero channels up to a multiple of 4 (34 -> 36): the pixel stride
    stays a multiple of 16 bytes, which the TMA loads of the tensor-core convolution need; conv_cl pads the
    weight's input channels to match."""
    C = a.shape[-1] + b.shape[-1]
    Cs = (C + pad_channels_to - 1) // pad_channels_to * pad_channels_to
    return _UpCatFn.apply(a, b, Cs)


def _strides3(rows, cols, trans=False):
    return (ctypes.c_longlong * 3)(0, 1, cols) if trans else (ctypes.c_longlong * 3)(0, cols, 1)


def gemm(A, B, C, M, N, K, sA, sB, sC, bias=None, batch=1, alpha=1.0, accumulate=False, relu=False):
    arr = lambda s: (ctypes.c_longlong * 3)(*[int(v) for v in s])
    _lib.call("dfmir_gemm", A, B, bias, C, batch, M, N, K, arr(sA), arr(sB), arr(sC), float(alpha), int(accumulate), int(relu))


class _LinearFn(torch.autograd.Function):
    """y = relu?(x W^T + b), x (M,K), W (N,K)  — nn.Linear of PatchSampleF.create_mlp."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        _lib.require_cuda(x, W)
        x, W = _f32(x).contiguous(), _f32(W).contiguous()
        M, K = x.shape
        N = W.shape[0]
        y = torch.empty((M, N), dtype=x.dtype, device=x.device)
        gemm(x, W, y, M, N, K, (0, K, 1), (0, 1, K), (0, N, 1), bias=b, relu=relu)
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.meta = (M, N, K, relu, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        M, N, K, relu, has_b = ctx.meta
        dy = _f32(dy).contiguous()
        if relu:
            g = torch.empty_like(dy)
            _lib.call("dfmir_act_bwd", y, dy, g, _lib.i64(y.numel()), ACT_RELU)
            dy = g
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            gemm(dy, W, dx, M, K, N, (0, N, 1), (0, K, 1), (0, K, 1))
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)   # dW (N,K) = dy^T x
            gemm(dy, x, dW, N, K, M, (0, 1, N), (0, K, 1), (0, K, 1))
        if has_b and ctx.needs_input_grad[2]:
            ones = torch.ones(M, dtype=dy.dtype, device=dy.device)
            db = torch.empty(N, dtype=dy.dtype, device=dy.device)
            gemm(ones, dy, db, 1, N, M, (0, M, 1), (0, N, 1), (0, N, 1))
        return dx, dW, db, None


def _split3_rows(x, b_style):
    rows, D = x.shape
    out = torch.empty((3 * rows, D), dtype=x.dtype, device=x.device)
    _lib.call("dfmir_tf32_split3_rows", x, out, _lib.i64(rows), D, int(b_style))
    return out


class _LinearTCFn(torch.autograd.Function):
    """nn.Linear on the tensor cores at fp32-class accuracy (the reference's nn.Linear is plain fp32: torch.matmul does
    not use TF32 by default): every product runs on the tcgen05 kernels with 3xTF32-split operands (hi*hi + hi*lo +
    lo*hi, models/networks.py:587-595 create_mlp).  A row-major (M, K) x (N, K)^T product IS a 1x1 convolution over M
    pixels: forward and dx on conv_umma_kernel (split along the reduction = channel axis), dW on the split-K
    weight-gradient kernel (split along the reduction = pixel axis)."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        _lib.require_cuda(x, W)
        x, W = _f32(x).contiguous(), _f32(W).contiguous()
        M, K = x.shape
        N = W.shape[0]
        y = torch.empty((M, N), dtype=x.dtype, device=x.device)
        xs, Ws = _split3(x, False), _split3(W, True)
        d = _make_desc(2, 1, 3 * K, N, [M // 64, 64], [M // 64, 64], [1, 1], [0, 0], 1, ACT_RELU if relu else ACT_NONE,
                       [M * 3 * K, 64 * 3 * K, 3 * K, 1], [M * N, 64 * N, N, 1])
        _run(lambda: _lib.call("dfmir_conv_umma_fwd", xs, Ws, b, y, ctypes.byref(d)), 2.0 * M * N * K, "umma_fwd", _nbytes(xs, Ws, y))
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.meta = (M, N, K, relu, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        M, N, K, relu, has_b = ctx.meta
        dy = _f32(dy).contiguous()
        if relu:
            g = torch.empty_like(dy)
            _lib.call("dfmir_act_bwd", y, dy, g, _lib.i64(y.numel()), ACT_RELU)
            dy = g
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            dys, WTs = _split3(dy, False), _split3(W.t().contiguous(), True)        # (M, 3N), (K, 3N)
            d = _make_desc(2, 1, 3 * N, K, [M // 64, 64], [M // 64, 64], [1, 1], [0, 0], 1, ACT_NONE,
                           [M * 3 * N, 64 * 3 * N, 3 * N, 1], [M * K, 64 * K, K, 1])
            _run(lambda: _lib.call("dfmir_conv_umma_fwd", dys, WTs, None, dx, ctypes.byref(d)), 2.0 * M * N * K, "umma_dgrad",
                 _nbytes(dys, WTs, dx))
        if ctx.needs_input_grad[1]:
            x3, dy3 = _split3_rows(x, False), _split3_rows(dy, True)                 # (3M, K), (3M, N)
            dw = torch.zeros((1, K, N), dtype=x.dtype, device=x.device)
            d = _make_desc(2, 1, K, N, [3 * M // 64, 64], [3 * M // 64, 64], [1, 1], [0, 0], 1, ACT_NONE,
                           [3 * M * K, 64 * K, K, 1], [3 * M * N, 64 * N, N, 1])
            _run(lambda: _lib.call("dfmir_conv_umma_wgrad", x3, dy3, dw, None, ctypes.byref(d)), 2.0 * M * N * K, "umma_wgrad",
                 _nbytes(x3, dy3, dw))
            dW = dw.view(K, N).t()
        if has_b and ctx.needs_input_grad[2]:
            ones = torch.ones(M, dtype=dy.dtype, device=dy.device)
            db = torch.empty(N, dtype=dy.dtype, device=dy.device)
            gemm(ones, dy, db, 1, N, M, (0, M, 1), (0, N, 1), (0, N, 1))
        return dx, dW, db, None


def linear(x, W, b, relu=False):
    """y = relu?(x W^T + b): large products (the PatchSampleF MLP at 4096+ rows) on the tensor cores with 3xTF32-split
    operands (_LinearTCFn), the rest on the fp32 GEMM kernel - fp32-class accuracy either way, like nn.Linear."""
    M, K = x.shape
    N = W.shape[0]
    if (CONV_ENGINE != "simt" and M >= UMMA_MIN_POSITIONS and M % 64 == 0 and K % 4 == 0 and K >= 16 and N % 4 == 0 and N >= 16
            and x.data_ptr() % 16 == 0):
        return _LinearTCFn.apply(x, W, b, relu)
    return _LinearFn.apply(x, W, b, relu)


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.require_cuda(x)
        x = _f32(x).contiguous()
        rows, D = x.shape
        y = torch.empty_like(x)
        norms = torch.empty(rows, dtype=x.dtype, device=x.device)
        _lib.call("dfmir_l2norm_fwd", x, y, norms, rows, D)
        ctx.save_for_backward(x, norms)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, norms = ctx.saved_tensors
        dy = _f32(dy).contiguous()
        dx = torch.empty_like(x)
        _lib.call("dfmir_l2norm_bwd", x, norms, dy, dx, x.shape[0], x.shape[1])
        return dx


def l2norm_rows(x):
    """x / (||x||_2 + 1e-7) per row (reference Normalize, models/networks.py:493-502)."""
    return _L2NormFn.apply(x)


class _GatherFn(torch.autograd.Function):
    """feat: logical (B,C,H,W) tensor of any strides; ids (P,) int64 -> (B*P, C)."""

    @staticmethod
    def forward(ctx, feat, ids):
        _lib.require_cuda(feat, ids)
        feat = _f32(feat)
        B, C, H, W = feat.shape
        P = ids.numel()
        ids = ids.to(torch.int64).contiguous()
        out = torch.empty((B * P, C), dtype=feat.dtype, device=feat.device)
        st = feat.stride()
        strides = (ctypes.c_longlong * 4)(st[0], st[2], st[3], st[1])
        _lib.call("dfmir_gather_patches_fwd", feat, ids, out, B, P, C, W, strides)
        ctx.save_for_backward(ids)
        ctx.meta = (B, C, H, W, P)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        B, C, H, W, P = ctx.meta
        dout = _f32(dout).contiguous()
        dfeat = torch.zeros((B, H, W, C), dtype=dout.dtype, device=dout.device)
        strides = (ctypes.c_longlong * 4)(H * W * C, W * C, C, 1)
        _lib.call("dfmir_gather_patches_bwd", dout, ids, dfeat, B, P, C, W, strides)
        return dfeat.permute(0, 3, 1, 2), None


class _GatherSparseFn(torch.autograd.Function):
    """gather_patches on a contiguous channels-last activation (B,H,W,C) whose gradient is returned as a hybrid
    sparse COO tensor (B*P rows of C values): the feature maps tapped for PatchNCE also feed the next layer, and
    autograd then adds the 256 patch rows per image into that layer's dense gradient (index_add) instead of
    materialising and adding a dense, almost entirely zero, gradient (537 MB at the generator's layer 4)."""

    @staticmethod
    def forward(ctx, x_cl, ids):
        _lib.require_cuda(x_cl, ids)
        B, H, W, C = x_cl.shape
        P = ids.numel()
        ids = ids.to(torch.int64).contiguous()
        out = torch.empty((B * P, C), dtype=x_cl.dtype, device=x_cl.device)
        strides = (ctypes.c_longlong * 4)(H * W * C, W * C, C, 1)
        _lib.call("dfmir_gather_patches_fwd", x_cl, ids, out, B, P, C, W, strides)
        ctx.save_for_backward(ids)
        ctx.meta = (B, C, H, W, P)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        B, C, H, W, P = ctx.meta
        b = torch.arange(B, device=ids.device).repeat_interleave(P)
        idx = torch.stack([b, (ids // W).repeat(B), (ids % W).repeat(B)])
        return torch.sparse_coo_tensor(idx, _f32(dout).reshape(B * P, C), (B, H, W, C), check_invariants=False), None


class _GatherSparsePadFn(torch.autograd.Function):
    """_GatherSparseFn for a tap that is the interior of a reflect-padded channels-last buffer P (B,H+2p,W+2p,C) - the
    ResnetBlock outputs: the gradient is a sparse COO tensor on P itself, so that autograd index-adds B*P rows into the
    dense gradient from the next block instead of zero-filling a dense (B,H,W,C) gradient, scattering into it, zero-filling
    a padded copy for the slice's backward and adding that (~2 GB of traffic per tapped layer and pass at batch 16)."""

    @staticmethod
    def forward(ctx, P_cl, ids, pad):
        _lib.require_cuda(P_cl, ids)
        B, HP, WP, C = P_cl.shape
        H, W = HP - 2 * pad, WP - 2 * pad
        n = ids.numel()
        ids = ids.to(torch.int64).contiguous()
        out = torch.empty((B * n, C), dtype=P_cl.dtype, device=P_cl.device)
        strides = (ctypes.c_longlong * 4)(HP * WP * C, WP * C, C, 1)
        _lib.call("dfmir_gather_patches_fwd", P_cl[:, pad:HP - pad, pad:WP - pad, :], ids, out, B, n, C, W, strides)
        ctx.save_for_backward(ids)
        ctx.meta = (B, C, HP, WP, W, n, pad)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        B, C, HP, WP, W, n, pad = ctx.meta
        b = torch.arange(B, device=ids.device).repeat_interleave(n)
        idx = torch.stack([b, (ids // W + pad).repeat(B), (ids % W + pad).repeat(B)])
        return torch.sparse_coo_tensor(idx, _f32(dout).reshape(B * n, C), (B, HP, WP, C), check_invariants=False), None, None


def gather_patches(feat, ids):
    """feat (B,C,H,W) logical layout -> (B*P, C) rows at the spatial positions ids."""
    pad = getattr(feat, "_dfmir_cl_pad", None)   # (P, p): feat is the interior of the padded buffer P (ResnetBlock outputs)
    if (SPARSE_TAP_GRAD and pad is not None and pad[0].requires_grad and torch.is_grad_enabled() and pad[0].dtype == torch.float32
            and pad[0].shape[3] == feat.shape[1] and pad[0].shape[1] == feat.shape[2] + 2 * pad[1]):
        return _GatherSparsePadFn.apply(pad[0], ids, pad[1])
    x_cl = getattr(feat, "_dfmir_cl", None)      # set by ResnetGenerator.forward on taps of plain channels-last outputs
    if (SPARSE_TAP_GRAD and x_cl is not None and x_cl.requires_grad and torch.is_grad_enabled() and x_cl.is_contiguous()
            and x_cl.dtype == torch.float32 and x_cl.shape == (feat.shape[0], feat.shape[2], feat.shape[3], feat.shape[1])):
        return _GatherSparseFn.apply(x_cl, ids)
    return _GatherFn.apply(feat, ids)


SPARSE_TAP_GRAD = os.environ.get("DFMIR_SPARSE_TAP_GRAD", "1") != "0"
if SPARSE_TAP_GRAD:
    torch.sparse.check_sparse_tensor_invariants.disable()     # explicit opt-out: indices come from our own randperm slices


def _split3(x, b_style):
    rows, D = x.shape
    out = torch.empty((rows, 3 * D), dtype=x.dtype, device=x.device)
    _lib.call("dfmir_tf32_split3", x, out, _lib.i64(rows), D, int(b_style))
    return out


class _PatchNCEFn(torch.autograd.Function):
    """PatchNCELoss.forward.  Tensor-core engine: S = Q K^T and dQ = dS K run on conv_umma_kernel as batched
    products of 3xTF32-split operands (hi*hi + hi*lo + lo*hi, fp32-class accuracy like the reference's fp32
    torch.bmm); fp32 engine: the CUDA-core GEMM."""

    @staticmethod
    def forward(ctx, q, k, B, T):
        _lib.require_cuda(q, k)
        q, k = _f32(q).contiguous(), _f32(k).contiguous()
        rows, D = q.shape
        if rows % B or k.shape != q.shape:
            raise _lib.DfmirError(f"PatchNCE: features {tuple(q.shape)} / {tuple(k.shape)} do not split into {B} images")
        P = rows // B
        S = torch.empty((B, P, P), dtype=q.dtype, device=q.device)
        loss = torch.empty(rows, dtype=q.dtype, device=q.device)
        tc = CONV_ENGINE != "simt" and P % 256 == 0 and D % 4 == 0 and D >= 16
        if tc:
            _lib.call("dfmir_patchnce_tc_fwd", _split3(q, False), _split3(k, True), S, loss, B, P, 3 * D, float(T))
        else:
            _lib.call("dfmir_patchnce_fwd", q, k, S, loss, B, P, D, float(T))
        ctx.save_for_backward(S, k)
        ctx.meta = (B, P, D, tc)
        return loss

    @staticmethod
    def backward(ctx, g):
        S, k = ctx.saved_tensors
        B, P, D, tc = ctx.meta
        g = _f32(g).contiguous()
        work = torch.empty_like(S)
        dq = torch.empty((B * P, D), dtype=g.dtype, device=g.device)
        if tc and D % 128 == 0:
            _lib.call("dfmir_patchnce_scale", S, g, work, B, P)
            kT = k.view(B, P, D).transpose(1, 2).contiguous().view(B * D, P)       # B[b]^T rows = feature dims
            _lib.call("dfmir_bmm_nt_umma", _split3(work.view(B * P, P), False), _split3(kT, True), dq, B, P, D, 3 * P)
        else:
            _lib.call("dfmir_patchnce_bwd", S, k, g, work, dq, B, P, D)
        return dq, None, None, None


def patchnce(q, k, batch, T):
    return _PatchNCEFn.apply(q, k, batch, T)
