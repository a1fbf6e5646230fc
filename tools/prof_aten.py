#!/usr/bin/env python
"""Which PyTorch (ATen) ops still launch kernels inside a step, by op and input shape (development aid)."""
import collections, contextlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dfmir_b200 import registration_model as rm
from torch.profiler import profile, ProfilerActivity
torch.cuda.set_device(0)
opt = rm.default_options(batch_size=16, crop_size=256, load_size=256, gpu_ids=[0])
torch.manual_seed(1234)
with contextlib.redirect_stdout(sys.stderr):
    model = rm.REGISTRATIONModel(opt)
    A, B = bench.synthetic_pair(16, 256, 1234)
    data = {"A": A.pin_memory(), "B": B.pin_memory()}
    model.data_dependent_initialize(data); model.setup(opt)
model.set_input(data)
for _ in range(3): model.optimize_parameters()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    model.optimize_parameters(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.key_averages(group_by_input_shape=True):
    if ev.key.startswith("aten::") and ev.self_device_time_total > 0:
        agg[(ev.key, str(ev.input_shapes)[:110])][0] += ev.self_device_time_total / 1e3
        agg[(ev.key, str(ev.input_shapes)[:110])][1] += ev.count
for (k, shp), (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{ms:8.3f} ms x{n:4d} {k:28s} {shp}")
