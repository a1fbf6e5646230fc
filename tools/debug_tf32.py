import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import numpy as np, torch, torch.nn.functional as F
import dfmir_b200.functional as Fn
from oracle import torch_port as tp
def rne(t):
    i = t.contiguous().view(torch.int32)
    lsb = (i >> 13) & 1
    i = (i + 0xFFF + lsb) & ~0x1FFF
    return i.view(torch.float32)
N, Cin, Cout, H, W, k, pad = 2, 128, 256, 16, 64, 3, 1
r = np.random.RandomState(1)
x = torch.from_numpy(r.standard_normal((N, Cin, H, W)).astype(np.float32))
w = torch.from_numpy((r.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32))
gy = torch.from_numpy(r.standard_normal((N, Cout, H, W)).astype(np.float32))
Fn.CONV_ENGINE = "auto"
xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
wg = w.cuda().requires_grad_()
yg = Fn.conv_cl(xg, wg, None, pad=pad)
yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
y_u = yg.detach().permute(0, 3, 1, 2).cpu().double(); dx_u = xg.grad.permute(0, 3, 1, 2).cpu().double(); dw_u = wg.grad.cpu().double()
for name, q in (("exact", lambda t: t), ("trunc", lambda t: tp.tf32_round(t, "trunc")), ("rna", lambda t: tp.tf32_round(t, "rna")), ("rne", rne)):
    xq, wq, gq = q(x).double(), q(w).double(), q(gy).double()
    y = F.conv2d(xq, wq, padding=pad)
    dx = torch.nn.grad.conv2d_input(x.shape, wq, gq, padding=pad)
    dw = torch.nn.grad.conv2d_weight(xq, w.shape, gq, padding=pad)
    print(f"{name:6s} fwd {float((y_u - y).abs().max() / y.abs().max()):.3e}  dgrad {float((dx_u - dx).abs().max() / dx.abs().max()):.3e}  wgrad {float((dw_u - dw).abs().max() / dw.abs().max()):.3e}")
