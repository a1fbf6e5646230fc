#!/usr/bin/env python
"""Per-kernel time breakdown (torch.profiler) of the 3-D bench step.
    python tools/breakdown3d.py                      # VoxelMorph-3D 128^3, batch 2, 6-level features (configs[2])
    python tools/breakdown3d.py 160,192,160 1 default  # configs[3]"""
import collections, os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from dfmir_b200 import vxm
from torch.profiler import profile, ProfilerActivity
shape = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (128, 128, 128)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
feats = None if len(sys.argv) > 3 and sys.argv[3] == "default" else [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]
torch.manual_seed(0)
net = vxm.VxmDense(shape, feats, int_steps=7, bidir=False).cuda()
opt = torch.optim.Adam(net.parameters(), lr=2e-4)
A = torch.rand(B, 1, *shape, device="cuda"); Bm = torch.rand(B, 1, *shape, device="cuda")
def step():
    opt.zero_grad()
    y, f, n, g = net.forward_with_losses(A, Bm)
    (n + 0.02 * g).backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        agg[ev.name][0] += ev.device_time / 1e3; agg[ev.name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"sum of kernel time {tot / 2:.2f} ms/step")
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{ms / 2:9.3f} ms/step {100 * ms / tot:6.2f}% x{n // 2:4d}  {name[:130]}")
