"""Diagnostic: generator encoder pass (fp32 engine) vs float64 torch port, dense synthetic loss and the PatchSampleF / PatchNCE loss, N = 1, 2."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from dfmir_b200 import networks
import dfmir_b200.functional as Fn
from dfmir_b200.patchnce import PatchNCELoss
from oracle import torch_port as tp
import argparse

sdG, sdF, _ = tp.random_state_dicts(ngf=64, n_blocks=9, crop=256, seed=5)
G = networks.define_G(1, 1, 64, 'resnet_9blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
G.load_state_dict(sdG, strict=False); G.cuda()
netF = networks.define_F(1, 'mlp_sample', 'instance', False, 'xavier', 0.02, False, [], argparse.Namespace(netF_nc=256))
ALL = [0, 4, 8, 12, 16]
layers = [int(v) for v in os.environ.get("TAPS", "0,4,8,12,16").split(",")]
Fn.SPARSE_TAP_GRAD = os.environ.get("SPARSE", "1") == "1"
S = 256
sizes = [(S + 6) ** 2, S * S, (S // 2) ** 2, (S // 4) ** 2, (S // 4) ** 2]
ids_all = [torch.from_numpy(np.random.RandomState(77 + i).permutation(n))[:256] for i, n in enumerate(sizes)]
ids = [ids_all[ALL.index(l)] for l in layers]
sdF = {k.replace(f"mlp_{ALL.index(l)}.", f"mlp_{j}."): v for j, l in enumerate(layers) for k, v in sdF.items() if k.startswith(f"mlp_{ALL.index(l)}.")}
Fn.CONV_ENGINE = sys.argv[1] if len(sys.argv) > 1 else "simt"
for N in (1,):
    x = torch.from_numpy(gi.image_textured(411, N, (256, 256)))
    k_img = torch.from_numpy(gi.image_textured(412, N, (256, 256)))
    for mode in os.environ.get("MODES", "nce").split(","):
        # oracle f64
        P = {k: v.double().clone().requires_grad_(not k.endswith('.filt')) for k, v in sdG.items()}
        PF = {k: v.double().clone().requires_grad_() for k, v in sdF.items()}
        xr = x.double().requires_grad_()
        feats = tp.resnet_generator(xr, P, 9, layers, encode_only=True)
        for f in feats: f.retain_grad()
        if mode == "dense":
            loss = sum((f * torch.from_numpy(gi.weights(500 + i, tuple(f.shape), 1.0)).double()).sum() / f.numel() ** 0.5 for i, f in enumerate(feats))
        elif mode in ("gather", "gathernorm"):
            rows, _ = tp.patch_sample(feats, None, 256, ids)
            if mode == "gather":
                rows = [f.permute(0, 2, 3, 1).flatten(1, 2)[:, ids[i], :].flatten(0, 1) for i, f in enumerate(feats)]
            loss = sum((r * torch.from_numpy(gi.weights(600 + i, tuple(r.shape), 1.0)).double()).sum() for i, r in enumerate(rows))
        else:
            with torch.no_grad():
                fk = tp.resnet_generator(k_img.double(), P, 9, layers, encode_only=True)
            kp, _ = tp.patch_sample(fk, PF, 256, ids)
            qp, _ = tp.patch_sample(feats, PF, 256, ids)
            loss = sum((tp.patchnce(q, k, N) * 0.25).mean() for q, k in zip(qp, kp)) / len(layers)
        loss.backward()
        # ours
        G.zero_grad()
        xg = x.cuda().requires_grad_()
        f2 = G(xg, layers, encode_only=True)
        for f in f2: f.retain_grad()
        if mode == "dense":
            l2 = sum((f * torch.from_numpy(gi.weights(500 + i, tuple(f.shape), 1.0)).cuda()).sum() / f.numel() ** 0.5 for i, f in enumerate(f2))
        elif mode in ("gather", "gathernorm"):
            rows2 = [Fn.gather_patches(f, ids[i].cuda()) for i, f in enumerate(f2)]
            if mode == "gathernorm":
                rows2 = [Fn.l2norm_rows(r) for r in rows2]
            l2 = sum((r * torch.from_numpy(gi.weights(600 + i, tuple(r.shape), 1.0)).cuda()).sum() for i, r in enumerate(rows2))
        else:
            if not netF.mlp_init:
                netF.create_mlp(f2); netF.load_state_dict(sdF); netF.cuda()
            netF.zero_grad()
            with torch.no_grad():
                fk2 = G(k_img.cuda(), layers, encode_only=True)
                kp2, _ = netF(fk2, 256, [t.cuda() for t in ids])
            qp2, _ = netF(f2, 256, [t.cuda() for t in ids])
            crit = PatchNCELoss(argparse.Namespace(batch_size=N, nce_T=0.07, nce_includes_all_negatives_from_minibatch=False))
            l2 = sum((crit(q, k) * 0.25).mean() for q, k in zip(qp2, kp2)) / len(layers)
        l2.backward()
        print(f"=== N={N} {mode} engine={Fn.CONV_ENGINE}: loss ours {float(l2):.6f} oracle {float(loss):.6f}")
        for k, p in G.named_parameters():
            if p.grad is None or not k.endswith("weight") or P[k].grad is None:
                continue
            g, ref = p.grad.cpu().double(), P[k].grad
            print(f"  {k:32s} |ref| {float(ref.norm()):9.3e} relerr {float((g - ref).norm() / ref.norm()):8.2e} cos {float((g * ref).sum() / (g.norm() * ref.norm())):.6f}")
        for i, (a, b) in enumerate(zip(f2, feats)):
            ga = a.grad.to_dense() if a.grad.is_sparse else a.grad
            print(f"  tap {layers[i]} grad: strides {tuple(a.grad.stride()) if not a.grad.is_sparse else 'sparse'} relerr {float((ga.cpu().double() - b.grad).norm() / b.grad.norm()):.2e}")
        g, ref = xg.grad.cpu().double(), xr.grad
        print(f"  {'dx':32s} |ref| {float(ref.norm()):9.3e} relerr {float((g - ref).norm() / ref.norm()):8.2e} cos {float((g * ref).sum() / (g.norm() * ref.norm())):.6f}")
        if mode == "nce":
            for k, p in netF.named_parameters():
                if k.endswith("0.weight"):
                    g, ref = p.grad.cpu().double(), PF[k].grad
                    print(f"  F.{k:30s} |ref| {float(ref.norm()):9.3e} relerr {float((g - ref).norm() / ref.norm()):8.2e}")
