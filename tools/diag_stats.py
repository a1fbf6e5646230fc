import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import dfmir_b200.functional as Fn
from dfmir_b200 import _lib
torch.manual_seed(0)
for (N, Cin, Cout, H, pad) in [(1, 128, 256, 128, 1), (1, 256, 256, 66, 0), (2, 64, 128, 256, 1), (1, 256, 128, 128, 1), (1, 128, 64, 256, 1)]:
    x = (torch.randn(N, H, H, Cin, device="cuda") * 0.7 + 0.3).relu()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (Cin * 9) ** 0.5
    b = torch.zeros(Cout, device="cuda")
    st = Fn.StatsSlot()
    y = Fn.conv_cl(x, w, b, pad=pad, stats_slot=st)
    assert st.rows is not None
    Ho = y.shape[1]
    ref_m = y.double().mean(dim=(1, 2)); ref_v = y.double().var(dim=(1, 2), unbiased=False)
    ref_r = 1.0 / torch.sqrt(ref_v + 1e-5)
    out = {}
    for name, slot in (("epilogue", st), ("pass", None)):
        stats_holder = {}
        a = Fn._InstNormFn.apply(y, None, True, 0, 0, 1e-5, None, None, slot)
        # statistics are saved for backward: recover them through the autograd context is awkward; recompute from a instead
        # a = relu((y - m) * r): use channels' positive entries to solve for r and m exactly is noisy; call the C entry directly
    N_, C_ = N, Cout
    stats1 = torch.empty((N, Cout, 2), device="cuda"); stats2 = torch.empty((N, Cout, 2), device="cuda")
    o1 = torch.empty_like(y); o2 = torch.empty_like(y)
    st2 = Fn.StatsSlot(); y2 = Fn.conv_cl(x, w, b, pad=pad, stats_slot=st2)
    _lib.call("dfmir_instnorm_fwd_rows", y2, None, o1, stats1, st2.rows, st2.rows.shape[1], N, Ho, Ho, Cout, 1e-5, 1, 0, 0)
    ws = Fn.workspace(_lib.lib().dfmir_instnorm_workspace_bytes(N, Cout), y.device)
    _lib.call("dfmir_instnorm_fwd", y2, None, o2, stats2, ws, _lib.size_t(ws.numel()), N, Ho, Ho, Cout, 1e-5, 1, 0, 0)
    for name, s in (("epilogue", stats1), ("pass", stats2)):
        m, r = s[..., 0].double(), s[..., 1].double()
        print(f"N={N} {Cin}->{Cout} @{H}: {name:9s} mean abs err / std {float(((m - ref_m).abs() * ref_r).max()):.2e}  rstd rel err {float(((r - ref_r).abs() / ref_r).max()):.2e}  |mean|/std max {float((ref_m.abs() * ref_r).max()):.2f}")
