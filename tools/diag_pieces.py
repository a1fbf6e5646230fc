import os, sys
import numpy as np, torch, torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
Fn.CONV_ENGINE = "simt"
torch.manual_seed(0)
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
for N in (1, 2):
    C, H, W = 16, 20, 20
    # instnorm with residual and halo
    for relu, pad, with_res in [(True, 1, False), (False, 1, True), (False, 0, True)]:
        x = torch.randn(N, C, H, W, dtype=torch.float64, requires_grad=True)
        res = torch.randn(N, C, H, W, dtype=torch.float64, requires_grad=True) if with_res else None
        y = F.instance_norm(x)
        if relu: y = F.relu(y)
        if with_res: y = y + res
        if pad: y = F.pad(y, (pad,) * 4, mode='reflect')
        gy = torch.randn_like(y)
        y.backward(gy)
        xg = x.detach().float().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        rg = None
        if with_res:
            rp = F.pad(res.detach().float(), (1,) * 4, mode='reflect').cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
            rg = rp
        yg = Fn.instnorm_cl(xg, relu=relu, out_pad=pad, res=rg, res_pad=1 if with_res else 0)
        yg.backward(gy.float().cuda().permute(0, 2, 3, 1).contiguous())
        msg = f"N={N} IN relu={relu} pad={pad} res={with_res}: fwd {rel(yg.detach().permute(0,3,1,2).cpu().double(), y.detach()):.2e} dx {rel(xg.grad.permute(0,3,1,2).cpu().double(), x.grad):.2e}"
        if with_res:
            msg += f" dres(interior) {rel(rg.grad[:, 1:-1, 1:-1, :].permute(0,3,1,2).cpu().double(), res.grad):.2e}"
        print(msg)
    # conv simt
    x = torch.randn(N, C, H + 2, W + 2, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(C, C, 3, 3, dtype=torch.float64) * 0.1).requires_grad_()
    b = torch.randn(C, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, b); gy = torch.randn_like(y); y.backward(gy)
    xg = x.detach().float().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    wg, bg = w.detach().float().cuda().requires_grad_(), b.detach().float().cuda().requires_grad_()
    yg = Fn.conv_cl(xg, wg, bg); yg.backward(gy.float().cuda().permute(0, 2, 3, 1).contiguous())
    print(f"N={N} conv: fwd {rel(yg.detach().permute(0,3,1,2).cpu().double(), y.detach()):.2e} dx {rel(xg.grad.permute(0,3,1,2).cpu().double(), x.grad):.2e} dw {rel(wg.grad.cpu().double(), w.grad):.2e} db {rel(bg.grad.cpu().double(), b.grad):.2e}")
    # pad
    x = torch.randn(N, C, H, W, dtype=torch.float64, requires_grad=True)
    y = F.pad(x, (1,) * 4, mode='reflect'); gy = torch.randn_like(y); y.backward(gy)
    xg = x.detach().float().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    yg = Fn.pad_reflect_cl(xg, 1); yg.backward(gy.float().cuda().permute(0, 2, 3, 1).contiguous())
    print(f"N={N} pad: dx {rel(xg.grad.permute(0,3,1,2).cpu().double(), x.grad):.2e}")
