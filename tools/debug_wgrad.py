import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch, torch.nn.functional as F
import dfmir_b200.functional as Fn
CASES = [(2, 64, 128, 16, 128, 3, 1), (2, 128, 256, 8, 64, 3, 1), (1, 256, 256, 18, 18, 3, 0), (2, 256, 128, 12, 40, 3, 1),
         (1, 128, 64, 9, 256, 3, 1), (1, 256, 256, 66, 66, 3, 0), (1, 128, 128, 8, 32, 1, 0), (1, 128, 128, 1, 32, 1, 0)]
for (N, Cin, Cout, H, W, k, pad) in CASES:
    r = np.random.RandomState(1)
    x = torch.from_numpy(r.standard_normal((N, Cin, H, W)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    y = F.conv2d(x, w, b, padding=pad)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    res = {}
    for eng in ("simt", "auto"):
        Fn.CONV_ENGINE = eng
        xg = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        yg = Fn.conv_cl(xg, wg, bg, pad=pad)
        yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
        torch.cuda.synchronize()
        res[eng] = (wg.grad.cpu(), bg.grad.cpu())
    sc = float(w.grad.abs().max())
    e_s = float((res["simt"][0] - w.grad).abs().max()) / sc
    e_u = float((res["auto"][0] - w.grad).abs().max()) / sc
    d = (res["auto"][0] - w.grad).abs()
    # per-tap error
    pt = d.amax(dim=(0, 1)).numpy() / sc
    eb = float((res["auto"][1] - b.grad).abs().max()) / float(b.grad.abs().max())
    ratio = float((res["auto"][0] * w.grad).sum() / (w.grad * w.grad).sum())
    print((N, Cin, Cout, H, W, k, pad), f"simt {e_s:.2e} umma {e_u:.2e} bias {eb:.2e} proj {ratio:.4f}")
    print("   per-tap rel err:", np.array2string(pt, precision=3))
    dd = d.amax(dim=(2, 3)).numpy() / sc     # (Cout, Cin)
    print("   by cout block of 32:", np.array2string(dd.reshape(Cout // 32, 32, Cin).max(axis=(1, 2)), precision=2),
          " by cin block:", np.array2string(dd.reshape(Cout, Cin // 32, 32).max(axis=(0, 2)), precision=2))
