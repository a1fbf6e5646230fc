import os, sys
import numpy as np, torch
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R_); sys.path.insert(0, os.path.join(R_, "tests", "golden"))
import inputs as gi
from oracle import torch_port as tp
from dfmir_b200 import vxm
import dfmir_b200.functional as Fn
S = 256
sds = tp.random_state_dicts(ngf=64, n_blocks=9, crop=S, seed=5)
sdR = sds[2]; sdR['flow.weight'] = sdR['flow.weight'] * 2e4
feats = [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]
R = vxm.VxmDense((S, S), feats, int_steps=7, bidir=True).cuda()
R.load_state_dict(sdR, strict=False)
for B in (1, 2):
    A = torch.from_numpy(gi.image_textured(700 + B, B, (S, S))).cuda()
    Bm = torch.from_numpy(gi.image_textured(710 + B, B, (S, S))).cuda()
    w1 = torch.from_numpy(gi.weights(1, (B, 1, S, S), 1.0)).cuda(); w2 = torch.from_numpy(gi.weights(2, (B, 2, S, S), 0.1)).cuda()
    res = {}
    for eng in ("simt", "auto"):
        Fn.CONV_ENGINE = eng
        R.zero_grad()
        prof = Fn.ConvProfile(); Fn.PROFILE = prof
        ys, yt, flow = R(A, Bm)
        ((ys * w1).sum() + (flow * w2).sum()).backward()
        Fn.PROFILE = None
        res[eng] = (ys.detach().clone(), flow.detach().clone(), {k: p.grad.clone() for k, p in R.named_parameters()}, prof.by_kind())
    s, a = res["simt"], res["auto"]
    print(f"B={B}: y_source relerr {float((a[0]-s[0]).norm()/s[0].norm()):.2e} flow relerr {float((a[1]-s[1]).norm()/s[1].norm()):.2e} |flow| max {float(s[1].abs().max()):.3f} kinds {{k: v[2] for k, v in a[3].items()}}".replace("{k", "").replace("}}", ""), {k: v[2] for k, v in a[3].items()})
    for k in s[2]:
        if k.endswith("weight"):
            print(f"   {k:40s} relerr {float((a[2][k]-s[2][k]).norm()/s[2][k].norm()):.2e}")
