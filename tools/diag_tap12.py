import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from dfmir_b200 import networks
import dfmir_b200.functional as Fn
from oracle import torch_port as tp
Fn.CONV_ENGINE = "simt"
sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=9, crop=256, seed=5)
G = networks.define_G(1, 1, 64, 'resnet_9blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
G.load_state_dict(sdG, strict=False); G.cuda()
m = G.model
ids = torch.from_numpy(np.random.RandomState(80).permutation(4096))[:256]
for N in (1, 2):
    x = torch.from_numpy(gi.image_textured(411, N, (256, 256)))
    w = torch.from_numpy(gi.weights(603, (N * 256, 256), 1.0))
    for mode in ("gather", "dense_same_values"):
        G.zero_grad()
        xg = x.cuda().requires_grad_()
        a = Fn.pad_reflect_cl(xg.permute(0, 2, 3, 1), 3)
        a = Fn.instnorm_cl(Fn.conv_cl(a, m[1].weight, m[1].bias), relu=True)
        a = Fn.blur_down_cl(Fn.instnorm_cl(Fn.conv_cl(a, m[4].weight, m[4].bias, pad=1), relu=True))
        a = Fn.blur_down_cl(Fn.instnorm_cl(Fn.conv_cl(a, m[8].weight, m[8].bias, pad=1), relu=True))
        Pin = Fn.pad_reflect_cl(a, 1); Pin.retain_grad()
        P = m[12].forward_padded(Pin, 1); P.retain_grad()
        v = P[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
        if mode == "gather":
            loss = (Fn.gather_patches(v, ids.cuda()) * w.cuda()).sum()
        else:
            dw = torch.zeros(N, 4096, 256); dw[:, ids, :] = w.view(N, 256, 256)
            loss = (v * dw.view(N, 64, 64, 256).permute(0, 3, 1, 2).cuda()).sum()
        loss.backward()
        g = P.grad
        halo = g.clone(); halo[:, 1:-1, 1:-1, :] = 0
        print(f"N={N} {mode}: P.grad strides {tuple(g.stride())} halo max {float(halo.abs().max()):.3e} interior nnz rows {int((g[:, 1:-1, 1:-1, :].abs().sum(-1) > 0).sum())}"
              f" | Pin.grad norm {float(Pin.grad.norm()):.6e} w12.1 grad norm {float(m[12].conv_block[1].weight.grad.norm()):.6e} w4 grad norm {float(m[4].weight.grad.norm()):.6e} dx norm {float(xg.grad.norm()):.6e}")

print("---- oracle (torch port) on CPU f64 / CUDA f64 / CUDA f32 vs ours")
for N in (1, 2):
    x = torch.from_numpy(gi.image_textured(411, N, (256, 256)))
    w = torch.from_numpy(gi.weights(603, (N * 256, 256), 1.0))
    G.zero_grad()
    xg = x.cuda().requires_grad_()
    f = G(xg, [12], encode_only=True)[0]
    (Fn.gather_patches(f, ids.cuda()) * w.cuda()).sum().backward()
    ours = {k: p.grad.detach().cpu().double() for k, p in G.named_parameters() if p.grad is not None and k.endswith("weight")}
    for dev, dt in (("cpu", torch.float64), ("cuda", torch.float64), ("cuda", torch.float32)):
        P = {k: v.to(dev).to(dt).clone().requires_grad_(not k.endswith('.filt')) for k, v in sdG.items()}
        xr = x.to(dev).to(dt).requires_grad_()
        fo = tp.resnet_generator(xr, P, 9, [12], encode_only=True)[0]
        rows = fo.permute(0, 2, 3, 1).flatten(1, 2)[:, ids.to(dev), :].flatten(0, 1)
        (rows * w.to(dev).to(dt)).sum().backward()
        line = f"N={N} oracle {dev}/{str(dt)[6:]}:"
        for k in ("model.4.weight", "model.12.conv_block.1.weight", "model.12.conv_block.5.weight"):
            ref = P[k].grad.detach().cpu().double()
            line += f" {k.split('.')[1]}{k.split('.')[-2] if 'block' in k else ''} |ref| {float(ref.norm()):.4e} relerr {float((ours[k] - ref).norm() / ref.norm()):.2e};"
        print(line)
