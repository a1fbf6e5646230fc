#!/usr/bin/env python
"""Times the fused cooperative launch (integrate -> resize -> warp -> NCC + Grad) against the chain of stand-alone
kernels at the 3-D BASELINE sizes, CUDA events, L2 flushed between repetitions; prints algorithmic GB/s against
the measured HBM peak (4*nd*N/2^nd velocity + 8N moving,fixed + 4N warped + 4*nd*N flow bytes per pair)."""
import json, os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tests", "golden"))
from dfmir_b200 import integrate_warp_loss, _lib
from test_gpu_fused import unfused
peak = json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(R, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, half in ((2, (64, 64, 64)), (1, (80, 96, 80)), (16, (128, 128))):
    nd = len(half); full = tuple(2 * s for s in half)
    N = 1
    for s in full: N *= s
    vel = torch.randn(B, nd, *half, device="cuda") * 2
    mov = torch.rand(B, 1, *full, device="cuda"); fix = torch.rand(B, 1, *full, device="cuda")
    alg = B * (4 * nd * N / 2 ** nd + 8 * N + 4 * N + 4 * nd * N)
    res = {}
    for name, fn in (("fused", lambda: integrate_warp_loss(vel, mov, fix, 7, 9)), ("unfused", lambda: unfused(vel, mov, fix, nd, 7, 9))):
        for _ in range(3): fn()
        ts = []; launches = 0
        for _ in range(10):
            flush.zero_(); torch.cuda.synchronize()
            _lib.launch_count_reset()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e)); launches = _lib.launch_count()
        ts.sort(); res[name] = (ts[len(ts) // 2], launches)
    f, u = res["fused"], res["unfused"]
    print(f"B={B} full={full}: fused {f[0]*1e3:8.1f} us ({f[1]} launch) = {alg/f[0]/1e6:7.1f} GB/s algorithmic = {alg/f[0]/1e6/peak:5.1%} of {peak:.0f} GB/s measured HBM"
          f" | unfused {u[0]*1e3:8.1f} us ({u[1]} launches) | speed-up {u[0]/f[0]:.2f}x | {B/(f[0]/1e3):.0f} pairs/s fwd")
