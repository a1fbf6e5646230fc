"""tcgen05 (UMMA) convolution engine bindings: which layer shapes run on the tensor cores (csrc/conv_umma.cu:
forward / data gradient; csrc/conv_umma_wgrad.cu: weight gradient) and the calls into libdfmir_b200.so for them.
Everything else (strided, one-channel and tiny layers) runs on the fp32 CUDA-core path (csrc/conv_simt.cu,
csrc/conv_thin.cu)."""
import ctypes
import weakref

from . import _lib

MIN_POSITIONS = 4096      # below this many output positions a layer is launch-bound: keep it on the fp32 kernels


def supported(d, dgrad=False):
    """This product (forward, or data gradient) of the convolution described by `d` fits the tcgen05 kernel
    (csrc/conv_umma.cu: 2-D / 3-D, stride 1, reduction-side channels a multiple of 4 and >= 16, channels-last
    operands with 16-byte aligned strides)."""
    return bool(_lib.lib().dfmir_conv_umma_supported(ctypes.byref(d), int(dgrad)))


def _run(fn, flops, kind, nbytes=0.0):
    from . import functional as Fn
    Fn._run(fn, flops, kind, nbytes)


def _nbytes(*tensors):
    return 4.0 * sum(t.numel() for t in tensors if t is not None)


_kmajor_cache = {}


def _kmajor(w):
    """[tap][Cout][Cin] copy (K-major rows for the B operand) of a packed weight; cached per packed tensor, which
    functional.packed_weight shares between the passes of a step."""
    key = id(w)
    hit = _kmajor_cache.get(key)
    if hit is not None and hit[0]() is w and hit[1] == w._version:
        return hit[2]
    if len(_kmajor_cache) > 256:
        for k in [k for k, v in _kmajor_cache.items() if v[0]() is None]:
            del _kmajor_cache[k]
    wk = w.detach().transpose(1, 2).contiguous()
    _kmajor_cache[key] = (weakref.ref(w), w._version, wk)
    return wk


def conv_fwd(x, w, bias, y, d, flops=0.0, stat_rows=None):
    wk = _kmajor(w)
    if stat_rows is not None:       # InstanceNorm statistics of y as a by-product of the epilogue
        _run(lambda: _lib.call("dfmir_conv_umma_fwd_stats", x, wk, bias, y, ctypes.byref(d), stat_rows), flops, "umma_fwd",
             _nbytes(x, wk, y))
        return
    _run(lambda: _lib.call("dfmir_conv_umma_fwd", x, wk, bias, y, ctypes.byref(d)), flops, "umma_fwd", _nbytes(x, wk, y))


def conv_dgrad(dy, w, dx, d, flops=0.0):
    # [tap][Cin][Cout] is K-major for this product
    _run(lambda: _lib.call("dfmir_conv_umma_dgrad", dy, w, dx, ctypes.byref(d)), flops, "umma_dgrad", _nbytes(dy, w, dx))


def conv_wgrad(x, dy, dw, db, d, flops=0.0):
    if _lib.lib().dfmir_conv_umma_wgrad_supported(ctypes.byref(d)):
        _run(lambda: _lib.call("dfmir_conv_umma_wgrad", x, dy, dw, db, ctypes.byref(d)), flops, "umma_wgrad", _nbytes(x, dy, dw))
    else:
        _run(lambda: _lib.call("dfmir_conv_wgrad", x, dy, dw, db, ctypes.byref(d)), flops, "simt")
