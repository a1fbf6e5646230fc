#!/usr/bin/env python
"""Pure kernel times (torch.profiler / CUPTI) of the registration kernels at 3-D 128^3 batch 2 and 2-D 256^2 batch 16."""
import os, sys, collections
import torch
from torch.profiler import profile, ProfilerActivity
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from dfmir_b200 import layers, losses, integrate_warp_loss
for B, half in ((2, (64, 64, 64)), (16, (128, 128))):
    nd = len(half); full = tuple(2 * s for s in half)
    vel = (torch.randn(B, nd, *half, device="cuda") * 2).requires_grad_()
    mov = torch.rand(B, 1, *full, device="cuda"); fix = torch.rand(B, 1, *full, device="cuda")
    vi = layers.VecInt(list(half), 7).cuda(); rs = layers.ResizeTransform(0.5, nd); st = layers.SpatialTransformer(list(full)).cuda()
    ncc = losses.NCC_Loss('cuda', kernel_var=[9] * nd); gl = losses.Grad_Loss(dim=nd)
    def step():
        flow = rs(vi(vel)); warped = st(mov, flow)
        (ncc(warped, fix) + gl(flow)).backward()
        w, f, n, g = integrate_warp_loss(vel.detach(), mov, fix, 7, 9)
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += ev.device_time; agg[ev.name][1] += 1
    print(f"--- B={B} full={full} (fwd + bwd of the unfused chain, then the fused forward)")
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"{us / 3:10.1f} us/iter x{n // 3:3d}  {name[:120]}")
