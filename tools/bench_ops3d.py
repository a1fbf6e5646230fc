#!/usr/bin/env python
"""CUDA-event timings of the stand-alone registration kernels at 3-D BASELINE sizes (development aid)."""
import os, sys, json
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from dfmir_b200 import layers, losses, integrate_warp_loss
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for B, half in ((2, (64, 64, 64)), (16, (128, 128))):
    nd = len(half); full = tuple(2 * s for s in half)
    vel = torch.randn(B, nd, *half, device="cuda") * 2
    mov = torch.rand(B, 1, *full, device="cuda"); fix = torch.rand(B, 1, *full, device="cuda")
    vi = layers.VecInt(list(half), 7).cuda(); rs = layers.ResizeTransform(0.5, nd); st = layers.SpatialTransformer(list(full)).cuda()
    ncc = losses.NCC_Loss('cuda', kernel_var=[9] * nd); gl = losses.Grad_Loss(dim=nd)
    field = vi(vel); flow = rs(field); warped = st(mov, flow)
    print(f"B={B} full={full}: vecint {timeit(lambda: vi(vel)):.1f} us | resize {timeit(lambda: rs(field)):.1f} | warp {timeit(lambda: st(mov, flow)):.1f} | "
          f"ncc {timeit(lambda: ncc(warped, fix)):.1f} | grad {timeit(lambda: gl(flow)):.1f} | fused {timeit(lambda: integrate_warp_loss(vel, mov, fix, 7, 9)):.1f}")
