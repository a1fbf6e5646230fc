#!/usr/bin/env python
"""CUDA-graph capture of the training step (REGISTRATIONModel.capture_step): eager vs replayed step time and losses.
    python tools/try_graph.py [--batch 16] [--size 256] [--steps 10]"""
import argparse, contextlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import bench
    from dfmir_b200 import registration_model as rm
    torch.cuda.set_device(0)
    opt = rm.default_options(batch_size=args.batch, crop_size=args.size, load_size=args.size, gpu_ids=[0], cuda_graph=True)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        model = rm.REGISTRATIONModel(opt)
        A, B = bench.synthetic_pair(args.batch, args.size, 1234)
        data = {"A": A.pin_memory(), "B": B.pin_memory()}
        model.data_dependent_initialize(data)
        model.setup(opt)
    model.set_input(data)

    def timed(n):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            model.set_input(data)
            model.optimize_parameters()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    for _ in range(15):
        model.optimize_parameters()
    print("eager   %.2f ms/step" % timed(args.steps), model.get_current_losses(), flush=True)
    t0 = time.time()
    model.capture_step()
    print("capture %.2f s, %d launches per step" % (time.time() - t0, model.graph_launches_per_step), flush=True)
    for _ in range(3):
        model.optimize_parameters()
    print("graph   %.2f ms/step" % timed(args.steps), model.get_current_losses(), flush=True)
    for _ in range(20):
        model.optimize_parameters()
    print("after 20 more replays", model.get_current_losses(), flush=True)


if __name__ == "__main__":
    main()
