#!/usr/bin/env python
"""Device time of the NCCL kernels inside one training step (torch.profiler / CUPTI on rank 0), bucketed overlapped all-reduce:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_timing.py"""
import collections, contextlib, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dfmir_b200 import registration_model as rm
from torch.profiler import profile, ProfilerActivity
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opt = rm.default_options(batch_size=16, crop_size=256, load_size=256, gpu_ids=[local])
torch.manual_seed(1234)
with contextlib.redirect_stdout(sys.stderr):
    model = rm.REGISTRATIONModel(opt)
    A, B = bench.synthetic_pair(16, 256, 1234 + rank)
    data = {"A": A.pin_memory(), "B": B.pin_memory()}
    model.data_dependent_initialize(data); model.setup(opt); model.parallelize()
model.set_input(data)
for _ in range(4): model.optimize_parameters()
torch.cuda.synchronize(); dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(3): model.optimize_parameters()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): model.optimize_parameters()
    torch.cuda.synchronize()
if rank == 0:
    agg = collections.defaultdict(lambda: [0.0, 0])
    tot = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            tot += ev.device_time / 1e3
            if "nccl" in ev.name.lower():
                agg[ev.name[:90]][0] += ev.device_time / 1e3; agg[ev.name[:90]][1] += 1
    print(f"world {dist.get_world_size()}: eager step {ms:.2f} ms (CUDA events); sum of all kernel time {tot / 3:.2f} ms/step; buckets {len(model._buckets)}; gradient bytes {model._flat_grad.numel() * 4 / 1e6:.1f} MB")
    for k, (t, n) in agg.items():
        print(f"   {t / 3:7.3f} ms/step x{n // 3:3d}  {k}")
dist.barrier(); dist.destroy_process_group()
