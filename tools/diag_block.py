import os, sys, functools
import numpy as np, torch, torch.nn as nn, torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
from dfmir_b200 import networks
Fn.CONV_ENGINE = "simt"
torch.manual_seed(0)
C, H = 16, 20
norm = functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
blk = networks.ResnetBlock(C, 'reflect', norm, False, True).cuda()
P_ = 32
for N in (1, 2):
    x = torch.randn(N, C, H, H)
    ids = torch.randperm(H * H)[:P_]
    gw = torch.randn(N * P_, C)
    # reference (float64 CPU)
    xr = x.double().requires_grad_()
    c1, c2 = blk.conv_block[1], blk.conv_block[5]
    h = F.relu(F.instance_norm(F.conv2d(F.pad(xr, (1,) * 4, mode='reflect'), c1.weight.detach().cpu().double(), c1.bias.detach().cpu().double())))
    h = F.instance_norm(F.conv2d(F.pad(h, (1,) * 4, mode='reflect'), c2.weight.detach().cpu().double(), c2.bias.detach().cpu().double()))
    out = xr + h
    rows = out.permute(0, 2, 3, 1).flatten(1, 2)[:, ids, :].flatten(0, 1)
    (rows * gw.double()).sum().backward()
    for variant in ("view_dense", "noslots", "nobias", "nores"):
        xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        Pin = Fn.pad_reflect_cl(xg, 1)
        op = 0 if variant == "pad0_gather" else 1
        if variant == "view_dense":
            Pout = blk.forward_padded(Pin, op)
        else:
            s1 = None if variant in ("noslots", "nobias") else Fn.BiasGradSlot()
            s2 = None if variant in ("noslots", "nobias") else Fn.BiasGradSlot()
            rs = None if variant in ("noslots", "nores") else Fn.ResidualGradSlot()
            y = Fn.conv_cl(Pin, c1.weight, c1.bias, bias_slot=s1, res_slot=rs)
            P1 = Fn.instnorm_cl(y, relu=True, out_pad=1, bias_slot=s1)
            y = Fn.conv_cl(P1, c2.weight, c2.bias, bias_slot=s2)
            Pout = Fn.instnorm_cl(y, relu=False, out_pad=op, res=Pin, res_pad=1, bias_slot=s2, res_slot=rs)
        print("   fwd relerr", float((Pout[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).cpu().double() - out.detach()).norm() / out.detach().norm()))
        inter = Pout[:, 1:-1, 1:-1, :] if op else Pout
        if variant == "contig_gather":
            inter = inter.contiguous()
        v = inter.permute(0, 3, 1, 2)
        if variant == "view_dense":
            dense_w = torch.zeros(N, H * H, C)
            dense_w[:, ids, :] = gw.view(N, P_, C)
            loss = (v * dense_w.view(N, H, H, C).permute(0, 3, 1, 2).cuda()).sum()
        else:
            loss = (Fn.gather_patches(v, ids.cuda()) * gw.cuda()).sum()
        loss.backward()
        g = xg.grad.permute(0, 3, 1, 2).cpu().double()
        print(f"N={N} {variant:14s} dx relerr {float((g - xr.grad).norm() / xr.grad.norm()):.3e}")
