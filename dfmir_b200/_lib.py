"""ctypes binding of libdfmir_b200.so (the C ABI declared in include/dfmir_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
or a call fails, a DfmirError is raised.  torch is used only for device memory and streams.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfmir_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dfmir_b200.h")


class DfmirError(RuntimeError):
    pass


_lib = None


def declared_symbols():
    """Names of every function declared in include/dfmir_b200.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfmir_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DfmirError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C dfmir_b200/csrc`). dfmir_b200 has no fallback path.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dfmir_last_error.restype = ctypes.c_char_p
        _lib.dfmir_launch_count.restype = ctypes.c_longlong
        for name in declared_symbols():
            if name.endswith("_workspace_bytes"):
                getattr(_lib, name).restype = ctypes.c_size_t
    return _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _ints(seq):
    return (ctypes.c_int * len(seq))(*[int(s) for s in seq])


def check(rc, what):
    if rc != 0:
        raise DfmirError(f"{what} failed ({rc}): {lib().dfmir_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DfmirError(
                "dfmir_b200 ops run on CUDA tensors only (sm_100a kernels; there is no CPU path). "
                f"Got a tensor on {t.device}.")


def call(name, *args):
    """Call a C-ABI entry point; raises DfmirError with the library's message on failure.  The launch goes to the
    current stream of the device that owns the first tensor argument (with that device made current for the call if
    it is not already), so a model built with gpu_ids=[k] works whatever the caller's current device is."""
    fn = getattr(lib(), name)
    conv = []
    dev = None
    for a in args:
        if isinstance(a, torch.Tensor) or a is None:
            if dev is None and a is not None and a.is_cuda:
                dev = a.device
            conv.append(_ptr(a))
        elif isinstance(a, float):
            conv.append(ctypes.c_float(a))
        elif isinstance(a, (list, tuple)):
            conv.append(_ints(a))
        elif isinstance(a, bool):
            conv.append(ctypes.c_int(int(a)))
        elif isinstance(a, int):
            conv.append(ctypes.c_int(a))
        else:
            conv.append(a)  # already a ctypes object (c_size_t, c_longlong, ...)
    if dev is None or dev.index == torch.cuda.current_device():
        check(fn(*conv, _stream()), name)
    else:
        with torch.cuda.device(dev):
            check(fn(*conv, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), name)


def size_t(v):
    return ctypes.c_size_t(int(v))


def i64(v):
    return ctypes.c_longlong(int(v))


def launch_count():
    return int(lib().dfmir_launch_count())


def launch_count_reset():
    lib().dfmir_launch_count_reset()
