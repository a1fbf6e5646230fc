import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import numpy as np, torch
import inputs as gi
from oracle import torch_port as tp
from dfmir_b200 import networks
import dfmir_b200.functional as Fn
S = 64
sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=4, crop=S, seed=3)
x = torch.from_numpy(gi.image_textured(411, 2, (S, S)))
layers = list(range(0, 26))
def ref(dtype, mode):
    tp.TF32_EMULATION = mode
    out, feats = tp.resnet_generator(x.to(dtype), {k: v.to(dtype) for k, v in sdG.items()}, 4, layers)
    tp.TF32_EMULATION = None
    return feats
f64 = ref(torch.float64, None); ft = ref(torch.float64, "trunc"); ft32 = ref(torch.float32, "trunc")
Fn.CONV_ENGINE = "auto"
G = networks.define_G(1, 1, 64, 'resnet_4blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
G.load_state_dict(sdG, strict=False); G.cuda()
with torch.no_grad():
    _, fu = G(x.cuda(), layers)
Fn.CONV_ENGINE = "simt"
with torch.no_grad():
    _, fs = G(x.cuda(), layers)
for i, l in enumerate(layers):
    a = f64[i]; sc = float(a.abs().max())
    e = lambda t: float((t.double().cpu() - a).abs().max()) / sc
    et = float((fu[i].double().cpu() - ft[i]).abs().max()) / sc
    e32 = float((ft32[i].double() - ft[i]).abs().max()) / sc
    print(f"layer {l:2d} shape {tuple(a.shape)} scale {sc:8.3f}: simt-exact {e(fs[i]):.2e} umma-exact {e(fu[i]):.2e} emu-exact {e(ft[i]):.2e} | umma-emu {et:.2e} emu32-emu64 {e32:.2e}")
