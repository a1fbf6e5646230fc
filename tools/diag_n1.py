"""Diagnostic: generator encoder pass gradients, tcgen05 engine vs exact fp32 engine, at batch 1 and 2."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from dfmir_b200 import networks
import dfmir_b200.functional as Fn
from oracle import torch_port as tp

sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=9, crop=256, seed=5)
G = networks.define_G(1, 1, 64, 'resnet_9blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
G.load_state_dict(sdG, strict=False); G.cuda()
layers = [0, 4, 8, 12, 16]
for N in (1, 2):
    x = torch.from_numpy(gi.image_textured(411, N, (256, 256))).cuda()
    res = {}
    for eng, env in (("simt", {}), ("auto", {}), ("auto_nostats", {"stats": False}), ("auto_noslots", {"slots": False})):
        Fn.CONV_ENGINE = "simt" if eng == "simt" else "auto"
        Fn.STATS_IN_EPILOGUE = env.get("stats", True)
        G.zero_grad()
        xg = x.clone().requires_grad_()
        full = os.environ.get("FULL", "0") == "1"
        if full:
            out = G(xg)
            feats = [out]
        else:
            feats = G(xg, layers, encode_only=True)
        loss = 0
        for i, f in enumerate(feats):
            w = torch.from_numpy(gi.weights(500 + i, tuple(f.shape), 1.0)).cuda()
            loss = loss + (f * w).sum() / f.numel() ** 0.5
        loss.backward()
        res[eng] = ({k: p.grad.detach().clone() for k, p in G.named_parameters() if p.grad is not None}, xg.grad.clone())
    for eng in ("auto", "auto_nostats"):
        print(f"--- N={N} {eng} vs simt")
        for k, g in res["simt"][0].items():
            if not k.endswith("weight"):
                continue
            a = res[eng][0][k]
            sc = float(g.abs().max())
            cos = float((a * g).sum() / (a.norm() * g.norm() + 1e-30))
            print(f"{k:34s} scale {sc:9.3e} relerr {float((a - g).abs().max()) / sc:8.2e} cos {cos:.6f}")
        a, g = res[eng][1], res["simt"][1]
        print(f"{'dx':34s} scale {float(g.abs().max()):9.3e} relerr {float((a - g).abs().max()) / float(g.abs().max()):8.2e}")
