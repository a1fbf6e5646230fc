"""numpy/ctypes front-end of oracle/dfmir_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdfmir_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "dfmir_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_grad.restype = ctypes.c_float
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ints(s):
    return (ctypes.c_int * len(s))(*[int(v) for v in s])


def normalized_grid(flow):
    flow = _f(flow)
    out = np.empty_like(flow)
    lib().orc_normalized_grid(_p(flow), _p(out), flow.shape[0], flow.shape[1], _ints(flow.shape[2:]))
    return out


def warp(src, flow, mode="bilinear", rcp_mul=False, return_idx=False):
    src, flow = _f(src), _f(flow)
    out = np.empty_like(src)
    idx = np.empty(flow.shape, dtype=np.int32) if return_idx else None
    lib().orc_warp(_p(src), _p(flow), _p(out), _p(idx), src.shape[0], src.shape[1], flow.shape[1],
                   _ints(flow.shape[2:]), {"bilinear": 0, "nearest": 1}[mode], int(rcp_mul))
    return (out, idx) if return_idx else out


def vecint(vec, nsteps, rcp_mul=False):
    vec = _f(vec)
    out = np.empty_like(vec)
    lib().orc_vecint(_p(vec), _p(out), vec.shape[0], vec.shape[1], _ints(vec.shape[2:]), int(nsteps), int(rcp_mul))
    return out


def resize(x, out_shape, pre_mul=1.0, post_mul=1.0):
    x = _f(x)
    y = np.empty(x.shape[:2] + tuple(out_shape), dtype=np.float32)
    lib().orc_resize(_p(x), _p(y), x.shape[0] * x.shape[1], x.ndim - 2, _ints(x.shape[2:]), _ints(out_shape),
                     ctypes.c_float(pre_mul), ctypes.c_float(post_mul))
    return y


def resize_transform(x, vel_resize):
    """ResizeTransform(vel_resize, ndims).forward  (layers.py:85-97)."""
    factor = 1.0 / vel_resize
    if factor == 1:
        return _f(x)
    out_shape = [int(s * factor) for s in x.shape[2:]]
    return resize(x, out_shape, 1.0, factor) if factor < 1 else resize(x, out_shape, factor, 1.0)


def ncc(I, J, mask=None, win=9, eps=1e-5, reduction=0, return_cc=False):
    I, J = _f(I), _f(J)
    mask = None if mask is None else _f(np.broadcast_to(mask, I.shape))
    out = np.zeros(3, dtype=np.float32)
    cc = np.empty_like(I) if return_cc else None
    lib().orc_ncc(_p(I), _p(J), _p(mask), _p(out), _p(cc), I.shape[0], I.ndim - 2, _ints(I.shape[2:]), int(win),
                  ctypes.c_float(eps), int(reduction))
    return (out, cc) if return_cc else out


def grad_loss(x, penalty=2, loss_mult=1.0):
    x = _f(x)
    return float(lib().orc_grad(_p(x), x.shape[0] * x.shape[1], x.ndim - 2, _ints(x.shape[2:]), int(penalty),
                                ctypes.c_float(loss_mult)))


def l1_masked(a, b, mask=None, mu=None, mv=None, thr=-0.95):
    a, b = _f(a), _f(b)
    m8 = None if mask is None else np.ascontiguousarray(np.broadcast_to(mask, a.shape), dtype=np.uint8)
    mu = None if mu is None else _f(mu)
    mv = None if mv is None else _f(mv)
    out = np.zeros(2, dtype=np.float32)
    lib().orc_l1_masked(_p(a), _p(b), _p(m8), _p(mu), _p(mv), ctypes.c_float(thr), _p(out), ctypes.c_longlong(a.size))
    return out


# ---------------------------------------------------------------- network layers (planar NCHW / NCDHW)
def conv(x, w, b=None, stride=1, pad=0):
    x, w = _f(x), _f(w)
    nd = x.ndim - 2
    N, Cin = x.shape[:2]
    Cout = w.shape[0]
    O = [(x.shape[2 + a] + 2 * pad - w.shape[2 + a]) // stride + 1 for a in range(nd)]
    y = np.empty((N, Cout, *O), dtype=np.float32)
    b = None if b is None else _f(b)
    lib().orc_conv(_p(x), _p(w), _p(b), _p(y), N, Cin, Cout, nd, _ints(x.shape[2:]), _ints(w.shape[2:]), int(stride), int(pad))
    return y


def instnorm(x, eps=1e-5):
    x = _f(x)
    y = np.empty_like(x)
    lib().orc_instnorm(_p(x), _p(y), x.shape[0] * x.shape[1], ctypes.c_longlong(int(np.prod(x.shape[2:]))), ctypes.c_float(eps))
    return y


def pad_reflect(x, p):
    x = _f(x)
    N, C, H, W = x.shape
    y = np.empty((N, C, H + 2 * p, W + 2 * p), dtype=np.float32)
    lib().orc_pad_reflect(_p(x), _p(y), N * C, H, W, int(p))
    return y


def blur_down(x):
    x = _f(x)
    N, C, H, W = x.shape
    y = np.empty((N, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=np.float32)
    lib().orc_blur_down(_p(x), _p(y), N * C, H, W)
    return y


def blur_up(x):
    x = _f(x)
    N, C, H, W = x.shape
    y = np.empty((N, C, 2 * H, 2 * W), dtype=np.float32)
    lib().orc_blur_up(_p(x), _p(y), N * C, H, W)
    return y


def upsample_nn(x):
    x = _f(x)
    y = np.empty(x.shape[:2] + tuple(2 * s for s in x.shape[2:]), dtype=np.float32)
    lib().orc_upsample_nn(_p(x), _p(y), x.shape[0] * x.shape[1], x.ndim - 2, _ints(x.shape[2:]))
    return y


def patchnce(q, k, batch, T=0.07):
    q, k = _f(q), _f(k)
    rows, D = q.shape
    loss = np.empty(rows, dtype=np.float32)
    lib().orc_patchnce(_p(q), _p(k), _p(loss), int(batch), rows // int(batch), D, ctypes.c_float(T))
    return loss
