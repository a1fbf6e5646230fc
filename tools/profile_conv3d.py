#!/usr/bin/env python
"""A full-resolution VoxelMorph-3D layer (default 16 -> 16, 3x3x3, 128^3, batch 2) forward + backward a few times,
the last repetition between cudaProfilerStart/Stop: target of the ncu captures of the small-channel 3-D kernels.
    python tools/profile_conv3d.py [Cin] [Cout] [size] [batch]"""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(0)
x = torch.randn(B, S, S, S, Cin, device="cuda", requires_grad=True)
w = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda") * 0.1).requires_grad_()
b = torch.zeros(Cout, device="cuda", requires_grad=True)
g = torch.randn(B, S, S, S, Cout, device="cuda")
for it in range(4):
    if it == 3:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = Fn.conv_cl(x, w, b, pad=1)
    e[1].record()
    y.backward(g)
    e[2].record()
    torch.cuda.synchronize()
    fl = 2.0 * B * S ** 3 * Cin * Cout * 27
    print(f"rep {it}: fwd {e[0].elapsed_time(e[1]):.3f} ms ({fl / e[0].elapsed_time(e[1]) / 1e9:.1f} TF/s)  "
          f"bwd {e[1].elapsed_time(e[2]):.3f} ms ({2 * fl / e[1].elapsed_time(e[2]) / 1e9:.1f} TF/s)", flush=True)
    x.grad = None; w.grad = None; b.grad = None
torch.cuda.cudart().cudaProfilerStop()
