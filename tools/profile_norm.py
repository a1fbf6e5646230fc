#!/usr/bin/env python
"""InstanceNorm(+ReLU, reflected halo) forward + backward on a ResnetBlock-sized activation (batch 16, 64x64x256),
a blur-pool and the 7x7 stem/head kernels a few times: target of the ncu captures of the memory-bound kernels."""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
x = torch.randn(B, 64, 64, 256, device="cuda", requires_grad=True)
g = torch.randn(B, 66, 66, 256, device="cuda")
img = torch.randn(B, 262, 262, 1, device="cuda", requires_grad=True)
w1 = (torch.randn(64, 1, 7, 7, device="cuda") * 0.1).requires_grad_()
for it in range(3):
    y = Fn.instnorm_cl(x, relu=True, out_pad=1)
    y.backward(g)
    f = Fn.conv_cl(img, w1, None)
    f.sum().backward()
    torch.cuda.synchronize()
print("ok")
