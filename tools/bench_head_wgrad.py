#!/usr/bin/env python
"""Weight gradient of the 3-D flow head (16 -> 3 channels, gradient padded to 4) at 2 x 128^3: tensor-core kernel vs the
fp32 few-output-channel kernel (development aid)."""
import ctypes, os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
from dfmir_b200 import _lib
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Cp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S, B = 128, 2
x = torch.randn(B, S, S, S, Cin, device="cuda")
dy = torch.randn(B, S, S, S, Cp, device="cuda")
d = Fn._make_desc(3, B, Cin, Cp, [S] * 3, [S] * 3, [3] * 3, [1] * 3, 1, 0, Fn._cl_strides(x, 3), Fn._cl_strides(dy, 3))
res = {}
for name in ("dfmir_conv_umma_wgrad", "dfmir_conv_wgrad"):
    if name == "dfmir_conv_umma_wgrad" and not _lib.lib().dfmir_conv_umma_wgrad_supported(ctypes.byref(d)):
        print(name, "unsupported"); continue
    dw = torch.zeros(27, Cin, Cp, device="cuda")
    for _ in range(2):
        _lib.call(name, x, dy, dw, None, ctypes.byref(d))
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        _lib.call(name, x, dy, dw, None, ctypes.byref(d))
    e.record(); torch.cuda.synchronize()
    dw.zero_(); _lib.call(name, x, dy, dw, None, ctypes.byref(d)); res[name] = dw.clone()
    print(f"{name}: {s.elapsed_time(e) / 5:.3f} ms")
if len(res) == 2:
    a, b = res.values()
    print("rel diff", float((a - b).abs().max() / b.abs().max()))
