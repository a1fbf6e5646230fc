import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
torch.manual_seed(0)
for B in (1, 2, 3):
    for C in (1, 128):
        H = W = 64
        P = 256
        feat = torch.randn(B, H, W, C, device="cuda")
        ids = torch.randperm(H * W, device="cuda")[:P]
        W1 = torch.randn(256, C, device="cuda") * 0.1; b1 = torch.randn(256, device="cuda") * 0.1
        W2 = torch.randn(256, 256, device="cuda") * 0.1; b2 = torch.randn(256, device="cuda") * 0.1
        gw = torch.randn(B * P, 256, device="cuda")
        res = {}
        for mode in ("ours_sparse", "ours_dense", "torch"):
            x = feat.clone().requires_grad_()
            y = x * 1.0
            v = y.permute(0, 3, 1, 2)
            if mode == "torch":
                xs = v.permute(0, 2, 3, 1).flatten(1, 2)[:, ids, :].flatten(0, 1)
                h = torch.relu(torch.nn.functional.linear(xs, W1, b1))
                h = torch.nn.functional.linear(h, W2, b2)
                o = h / (h.pow(2).sum(1, keepdim=True).pow(0.5) + 1e-7)
            else:
                if mode == "ours_sparse":
                    v._dfmir_cl = y
                xs = Fn.gather_patches(v, ids)
                h = Fn.linear(xs, W1, b1, relu=True)
                h = Fn.linear(h, W2, b2)
                o = Fn.l2norm_rows(h)
            (o * gw).sum().backward()
            g = x.grad; res[mode] = (o.detach(), (g.to_dense() if g.is_sparse else g).clone())
        for m in ("ours_sparse", "ours_dense"):
            print(f"B={B} C={C} {m}: out err {float((res[m][0]-res['torch'][0]).abs().max()):.2e} grad relerr "
                  f"{float((res[m][1]-res['torch'][1]).norm()/res['torch'][1].norm()):.2e}")
