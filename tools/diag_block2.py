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
c1, c2 = blk.conv_block[1], blk.conv_block[5]
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
for N in [int(v) for v in os.environ.get("NS", "1,2").split(",")]:
    x = torch.randn(N, C, H, H)
    if os.environ.get("MODE", "") == "gather":
        ids = torch.randperm(H * H)[:32]
        rows_w = torch.randn(N * 32, C)
        dwn = torch.zeros(N, H * H, C); dwn[:, ids, :] = rows_w.view(N, 32, C)
        gw = dwn.view(N, H, H, C).permute(0, 3, 1, 2).contiguous()
    else:
        gw = torch.randn(N, C, H, H)
    if os.environ.get("SPARSE_SUPPORT", "0") == "1":
        ids = torch.randperm(H * H)[:32]
        m = torch.zeros(H * H); m[ids] = 1.0
        gw = gw * m.view(1, 1, H, H)
    xr = x.double().requires_grad_()
    Pin_r = F.pad(xr, (1,) * 4, mode='reflect'); Pin_r.retain_grad()
    y1_r = F.conv2d(Pin_r, c1.weight.detach().cpu().double(), c1.bias.detach().cpu().double()); y1_r.retain_grad()
    P1_r = F.pad(F.relu(F.instance_norm(y1_r)), (1,) * 4, mode='reflect'); P1_r.retain_grad()
    y2_r = F.conv2d(P1_r, c2.weight.detach().cpu().double(), c2.bias.detach().cpu().double()); y2_r.retain_grad()
    out_r = xr + F.instance_norm(y2_r)
    (out_r * gw.double()).sum().backward()
    xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    Pin = Fn.pad_reflect_cl(xg, 1); Pin.retain_grad()
    y1 = Fn.conv_cl(Pin, c1.weight, c1.bias); y1.retain_grad()
    P1 = Fn.instnorm_cl(y1, relu=True, out_pad=1); P1.retain_grad()
    y2 = Fn.conv_cl(P1, c2.weight, c2.bias); y2.retain_grad()
    OP = int(os.environ.get("OP", "0"))
    Pout = Fn.instnorm_cl(y2, relu=False, out_pad=OP, res=Pin, res_pad=1)
    if OP:
        full = Pout
        Pout = Pout[:, OP:-OP, OP:-OP, :]
        v = Pout.permute(0, 3, 1, 2)
        if os.environ.get("LOSS", "mul") == "mul":
            (v * gw.cuda()).sum().backward()
        else:
            rows = Fn.gather_patches(v, ids.cuda())
            (rows * rows_w.cuda()).sum().backward()
    else:
        (Pout * gw.cuda().permute(0, 2, 3, 1)).sum().backward()
    f = lambda t: t.permute(0, 3, 1, 2).cpu().double()
    print(f"N={N}: fwd y1 {rel(f(y1.detach()), y1_r.detach()):.1e} P1 {rel(f(P1.detach()), P1_r.detach()):.1e} y2 {rel(f(y2.detach()), y2_r.detach()):.1e} out {rel(f(Pout.detach()), out_r.detach()):.1e}")
    print(f"N={N}: grad y2 {rel(f(y2.grad), y2_r.grad):.1e} P1 {rel(f(P1.grad), P1_r.grad):.1e} y1 {rel(f(y1.grad), y1_r.grad):.1e} Pin {rel(f(Pin.grad), Pin_r.grad):.1e} x {rel(f(xg.grad), xr.grad):.1e}")
