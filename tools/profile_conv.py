#!/usr/bin/env python
"""Runs one ResnetBlock-shaped convolution (256 -> 256, 3x3 on a reflect-padded 66x66 buffer, batch 16) forward +
backward a few times on the tcgen05 engine: the target of the `ncu --set full` captures under profiles/.
Also prints CUDA-event timings of the three kernels (fwd, dgrad, wgrad)."""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
from dfmir_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
x = torch.randn(B, 66, 66, 256, device="cuda", requires_grad=True)
w = (torch.randn(256, 256, 3, 3, device="cuda") * 0.02).requires_grad_()
b = torch.zeros(256, device="cuda", requires_grad=True)
gy = torch.randn(B, 64, 64, 256, device="cuda")
flops = 2.0 * B * 64 * 64 * 256 * 2304
for it in range(reps):
    prof = Fn.ConvProfile(); Fn.PROFILE = prof
    y = Fn.conv_cl(x, w, b)
    y.backward(gy)
    Fn.PROFILE = None
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) for _, s, e, _ in prof.events]
    print("rep %d: fwd %.3f ms (%.0f TF/s)  dgrad %.3f ms (%.0f TF/s)  wgrad+bias %.3f ms (%.0f TF/s)" % (
        it, ts[0], flops / ts[0] / 1e9, ts[1], flops / ts[1] / 1e9, ts[2], flops / ts[2] / 1e9))
    x.grad = None; w.grad = None; b.grad = None
