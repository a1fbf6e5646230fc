#!/usr/bin/env python
"""Per-kernel time breakdown of one training step (torch.profiler / CUPTI; warm caches, low overhead).
Development aid: the committed evidence is the ncu launch list under profiles/.

    python tools/kernel_breakdown.py [--batch 16] [--size 256] [--steps 2] [--out gpurun_out/breakdown.txt]
"""
import argparse
import collections
import contextlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import bench
    from dfmir_b200 import registration_model as rm
    torch.cuda.set_device(0)
    opt = rm.default_options(batch_size=args.batch, crop_size=args.size, load_size=args.size, gpu_ids=[0])
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        model = rm.REGISTRATIONModel(opt)
        A, B = bench.synthetic_pair(args.batch, args.size, 1234)
        data = {"A": A.pin_memory(), "B": B.pin_memory()}
        model.data_dependent_initialize(data)
        model.setup(opt)
    model.set_input(data)
    for _ in range(3):
        model.optimize_parameters()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        model.optimize_parameters()
    e.record()
    torch.cuda.synchronize()
    wall = s.elapsed_time(e) / args.steps
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            model.optimize_parameters()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += ev.device_time / 1e3
            agg[ev.name][1] += 1
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    total = sum(v[0] for _, v in rows)
    lines = [f"step (CUDA events, unprofiled): {wall:.2f} ms; sum of kernel time under profiler: {total / args.steps:.2f} ms/step; "
             f"{sum(v[1] for _, v in rows) // args.steps} launches/step"]
    for name, (ms, n) in rows[:60]:
        lines.append(f"{ms / args.steps:10.3f} ms/step {100 * ms / total:6.2f}%  x{n // args.steps:5d}  {name[:150]}")
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
