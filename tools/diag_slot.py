"""Diagnostic: two chained ResnetBlocks, each engine vs float64 (plain / TF32-truncated operands): error structure."""
import os, sys, functools
import numpy as np, torch, torch.nn as nn, torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from oracle import torch_port as tp
import dfmir_b200.functional as Fn
from dfmir_b200 import networks

r = gi.rng(1300)
N, C, Hh = 2, 256, 24
norm = functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
blocks = [networks.ResnetBlock(C, 'reflect', norm, False, True) for _ in range(2)]
for b in blocks:
    for conv in (b.conv_block[1], b.conv_block[5]):
        conv.weight.data = torch.from_numpy((r.standard_normal((C, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32))
        conv.bias.data = torch.from_numpy(r.standard_normal(C).astype(np.float32) * 0.1)
x = torch.from_numpy(r.standard_normal((N, C, Hh, Hh)).astype(np.float32))
gy = torch.from_numpy(r.standard_normal((N, C, Hh, Hh)).astype(np.float32))


def ref(emulate, nblocks=2):
    tp.TF32_EMULATION = emulate
    try:
        xr = x.double().cuda().requires_grad_()
        a = xr
        inter = []
        for b in blocks[:nblocks]:
            w1, b1 = b.conv_block[1].weight.detach().double().cuda(), b.conv_block[1].bias.detach().double().cuda()
            w2, b2 = b.conv_block[5].weight.detach().double().cuda(), b.conv_block[5].bias.detach().double().cuda()
            y1 = tp.conv2d(F.pad(a, (1,) * 4, mode='reflect'), w1, b1); y1.retain_grad()
            h = F.relu(F.instance_norm(y1))
            y2 = tp.conv2d(F.pad(h, (1,) * 4, mode='reflect'), w2, b2); y2.retain_grad()
            h = F.instance_norm(y2)
            a = a + h
            inter += [y1, y2]
        (a * gy.double().cuda()).sum().backward()
    finally:
        tp.TF32_EMULATION = None
    return a.detach().cpu(), xr.grad.cpu(), [t.grad.cpu() for t in inter], [t.detach().cpu() for t in inter]


def run(engine, use_slots, stats=True, nblocks=2):
    Fn.CONV_ENGINE, Fn.UMMA_MIN_POSITIONS, Fn.STATS_IN_EPILOGUE = engine, 0, stats
    for b in blocks:
        b.cuda(); b.zero_grad()
    xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    P = Fn.pad_reflect_cl(xg, 1)
    for i, b in enumerate(blocks[:nblocks]):
        op = 1 if i < nblocks - 1 else 0
        if use_slots:
            P = b.forward_padded(P, op)
        else:
            c1, c2 = b.conv_block[1], b.conv_block[5]
            y = Fn.conv_cl(P, c1.weight, c1.bias)
            P1 = Fn.instnorm_cl(y, relu=True, out_pad=1)
            y = Fn.conv_cl(P1, c2.weight, c2.bias)
            P = Fn.instnorm_cl(y, relu=False, out_pad=op, res=P, res_pad=1)
    (P * gy.cuda().permute(0, 2, 3, 1)).sum().backward()
    Fn.CONV_ENGINE, Fn.UMMA_MIN_POSITIONS, Fn.STATS_IN_EPILOGUE = "auto", 4096, True
    return P.detach().permute(0, 3, 1, 2).cpu().double(), xg.grad.permute(0, 3, 1, 2).cpu().double()


def rel(a, b):
    d = (a - b).abs()
    i = int(d.argmax())
    idx = np.unravel_index(i, tuple(d.shape))
    return f"relnorm {float((a-b).norm()/b.norm()):.2e} max {float(d.max()):.2e} at {tuple(int(v) for v in idx)} (scale {float(b.abs().max()):.2f}) p99.9 {float(d.flatten().kthvalue(int(d.numel()*0.999))[0]):.2e}"


for nb in (1, 2):
    o64 = ref(None, nb); otr = ref("trunc", nb)
    print(f"--- {nb} block(s)")
    print("trunc-f64 vs f64:       out", rel(otr[0], o64[0]), "| dx", rel(otr[1], o64[1]))
    for eng, slots, stats in [("simt", False, True), ("simt", True, True), ("auto", False, True), ("auto", True, True), ("auto", False, False), ("auto", True, False)]:
        out, dx = run(eng, slots, stats, nb)
        want = o64 if eng == "simt" else otr
        print(f"{eng:5s} slots={int(slots)} stats_epi={int(stats)}: out", rel(out, want[0]), "| dx", rel(dx, want[1]))
