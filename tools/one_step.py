#!/usr/bin/env python
"""One training step of the bench workload between cudaProfilerStart/Stop (after initialisation and 3 warm-up
steps), for `ncu --profile-from-start off`: the launch list and the full captures under profiles/ come from it.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/one_step.py [--batch 16] [--size 256]
"""
import argparse
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
    torch.cuda.cudart().cudaProfilerStart()
    model.optimize_parameters()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("losses", model.get_current_losses())


if __name__ == "__main__":
    main()
