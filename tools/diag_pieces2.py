import os, sys
import numpy as np, torch, torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
Fn.CONV_ENGINE = "simt"
torch.manual_seed(0)
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
for N in (1, 2):
    C, H, W = 256, 64, 64
    for sparse in (False, True):
        ids = torch.randperm(H * W)[:256]
        msk = torch.zeros(H * W, dtype=torch.float64); msk[ids] = 1.0
        x = torch.randn(N, C, H, W, dtype=torch.float64, requires_grad=True)
        res = torch.randn(N, C, H, W, dtype=torch.float64, requires_grad=True)
        y = F.instance_norm(x) + res
        yp = F.pad(y, (1,) * 4, mode='reflect')
        gy = torch.randn_like(y)
        if sparse: gy = gy * msk.view(1, 1, H, W)
        gyp = F.pad(gy, (1,) * 4)          # zero halo gradient, as through an interior view
        yp.backward(gyp)
        xg = x.detach().float().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        rp = F.pad(res.detach().float(), (1,) * 4, mode='reflect').cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        slot = Fn.BiasGradSlot()
        yg = Fn.instnorm_cl(xg, relu=False, out_pad=1, res=rp, res_pad=1, bias_slot=slot)
        yg.backward(gyp.float().cuda().permute(0, 2, 3, 1).contiguous())
        print(f"N={N} sparse={sparse} IN: fwd {rel(yg.detach().permute(0,3,1,2).cpu().double(), yp.detach()):.2e} dx {rel(xg.grad.permute(0,3,1,2).cpu().double(), x.grad):.2e}"
              f" dres {rel(rp.grad[:, 1:-1, 1:-1, :].permute(0,3,1,2).cpu().double(), res.grad):.2e} dbias {rel(slot.db.cpu().double(), x.grad.sum(dim=(0,2,3))):.2e} (|db| {float(slot.db.abs().max()):.1e})")
        # conv wgrad / dgrad with this kind of dy
        xc = torch.randn(N, C, H + 2, W + 2, dtype=torch.float64, requires_grad=True)
        w = (torch.randn(C, C, 3, 3, dtype=torch.float64) * 0.05).requires_grad_()
        b = torch.zeros(C, dtype=torch.float64, requires_grad=True)
        yc = F.conv2d(xc, w, b)
        gc = x.grad.detach()               # the instance-norm backward's output as dy
        yc.backward(gc)
        xcg = xc.detach().float().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        wg, bg = w.detach().float().cuda().requires_grad_(), b.detach().float().cuda().requires_grad_()
        ycg = Fn.conv_cl(xcg, wg, bg); ycg.backward(gc.float().cuda().permute(0, 2, 3, 1).contiguous())
        print(f"N={N} sparse={sparse} conv: dx {rel(xcg.grad.permute(0,3,1,2).cpu().double(), xc.grad):.2e} dw {rel(wg.grad.cpu().double(), w.grad):.2e}")
