import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import numpy as np, torch
import inputs as gi
from oracle import torch_port as tp
from dfmir_b200 import networks
import dfmir_b200.functional as Fn
scale_w = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=4, crop=S, seed=3)
sdG = {k: (v * scale_w if k.endswith("weight") else v) for k, v in sdG.items()}
x = torch.from_numpy(gi.image_textured(411, 2, (S, S)))
wts = torch.from_numpy(gi.weights(412, (2, 1, S, S), 1.0))
def ref(dtype):
    leaves = {k: v.clone().to(dtype).requires_grad_() for k, v in sdG.items()}
    out = tp.resnet_generator(x.to(dtype), leaves, 4)
    (out * wts.to(dtype)).sum().backward()
    return out.detach(), {k: v.grad for k, v in leaves.items()}
o32, g32 = ref(torch.float32)
o64, g64 = ref(torch.float64)
emu = {}
for mode in ("trunc", "rna"):
    tp.TF32_EMULATION = mode
    emu[mode] = ref(torch.float64)
tp.TF32_EMULATION = None
res = {}
for eng in ("simt", "auto"):
    Fn.CONV_ENGINE = eng
    G = networks.define_G(1, 1, 64, 'resnet_4blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
    G.load_state_dict(sdG, strict=False); G.cuda()
    out = G(x.cuda()); (out * wts.cuda()).sum().backward()
    res[eng] = (out.detach().cpu(), {k: p.grad.cpu() for k, p in G.named_parameters()})
print("out: umma vs emu-trunc %.2e, vs emu-rna %.2e" % (float((res["auto"][0].double() - emu["trunc"][0]).abs().max()), float((res["auto"][0].double() - emu["rna"][0]).abs().max())))
print("out err: cpu32 %.2e simt %.2e umma %.2e" % tuple(float((o.double() - o64).abs().max()) for o in (o32, res["simt"][0], res["auto"][0])))
for k in g64:
    if k.endswith("weight"):
        sc = float(g64[k].abs().max())
        e = [float((g.double() - g64[k]).abs().max()) / sc for g in (g32[k], res["simt"][1][k], res["auto"][1][k])]
        et = float((res["auto"][1][k].double() - emu["trunc"][1][k]).abs().max()) / sc
        er = float((res["auto"][1][k].double() - emu["rna"][1][k]).abs().max()) / sc
        print(f"{k:32s} scale {sc:9.3e}  cpu32 {e[0]:.2e}  simt {e[1]:.2e}  umma {e[2]:.2e} | umma vs emu-trunc {et:.2e} vs emu-rna {er:.2e}")
