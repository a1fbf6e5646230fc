#!/usr/bin/env python
"""bench.py — volume-pairs/sec (fwd+bwd+Adam) of the DFMIR translation+registration training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

Workload (BASELINE.json configs[1]): 2-D 256x256, batch 16 MR->CT pairs per GPU, one full
REGISTRATIONModel.optimize_parameters (ResnetGenerator-9blocks x (2 full + 6 encoder passes),
VoxelMorph-2D + VecInt, PatchNCE x3, masked L1 x2, smoothing, backward, 3 Adam steps), synthetic
images, random-initialised weights.  One process per GPU (torchrun for N > 1, NCCL gradient
all-reduce); weak scaling (per-GPU batch fixed).  After the warm-up the step is captured into one CUDA graph
(REGISTRATIONModel.capture_step; DFMIR_CUDA_GRAPH=0 keeps eager launches) and K replays are timed with the inputs
resident (`value`) and through set_input + get_current_losses (`e2e`); K further eager steps with CUDA events
around every convolution launch give the `roofline` / `kernels` figures.  Prints ONE JSON line on rank 0.
`--workload 3d`: VoxelMorph-3D step (configs[2]; `--shape3d 160,192,160 --features3d default` for configs[3]).

`--impl reference` times the reference's CPU PyTorch path (oracle/torch_port.py, pinned to the
reference's own step by tests/test_oracle_nets.py) on the host cores, one pair per step.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "volume-pairs/sec (fwd+bwd)"
UNIT = "pairs/s"


def synthetic_pair(B, S, seed):
    """Smooth blob images in [-1, 1] with an exact -1 background outside a centred ellipse."""
    g = torch.Generator().manual_seed(seed)
    out = []
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    inside = (xx * xx + yy * yy) < 0.8
    for dom in range(2):
        x = torch.randn(B, 1, S, S, generator=g)
        k = torch.ones(1, 1, 9, 9) / 81.0
        for _ in range(3):
            x = torch.nn.functional.conv2d(x, k, padding=4)
        x = torch.tanh(3.0 * x / x.std()) * 0.9 + 0.03 * torch.randn(B, 1, S, S, generator=g)
        x = torch.where(inside[None, None], x.clamp(-0.94, 1.0), torch.full_like(x, -1.0))
        out.append(x.contiguous())
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, dev):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(dev), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        # median of the samples taken under load (upper half), as idle samples precede the region
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """The reference's CPU PyTorch path (oracle/torch_port.Step) on all host cores: one pair per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S = args.size
    G, Fd, R = tp.random_state_dicts(crop=S)
    A, B = synthetic_pair(1, S, 1234)
    st = tp.Step(G, Fd, R, n_blocks=9, batch_size=1, dvf_image=torch.zeros(1, 3, S, S))
    for _ in range(args.warmup):
        st.step(A, B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step(A, B)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} steps of 1 pair (batch 1) at {S}x{S}, torch CPU fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"2D {S}x{S} translation+registration fwd/bwd+Adam (BASELINE configs[1]); CPU sample: batch 1 per step",
                   "batch_per_step": 1},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(S, budget_s=25.0):
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G, Fd, R = tp.random_state_dicts(crop=S)
    A, B = synthetic_pair(1, S, 1234)
    st = tp.Step(G, Fd, R, n_blocks=9, batch_size=1, dvf_image=torch.zeros(1, 3, S, S))
    st.step(A, B)
    n, t0 = 0, time.perf_counter()
    while n < 2 or (time.perf_counter() - t0 < budget_s and n < 8):
        st.step(A, B)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} steps of 1 pair (batch 1) at {S}x{S} after 1 warm-up, oracle/torch_port.py (torch CPU fp32, {cores} threads)"}


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dfmir_b200 import _lib, registration_model as rm
    from dfmir_b200 import functional as Fn

    B, S = args.batch, args.size
    use_graph = os.environ.get("DFMIR_CUDA_GRAPH", "1") != "0"
    opt = rm.default_options(batch_size=B, crop_size=S, load_size=S, gpu_ids=[local], cuda_graph=use_graph)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(sys.stderr):
        model = rm.REGISTRATIONModel(opt)
        A, Bm = synthetic_pair(B, S, 1234 + rank)
        data = {"A": A.pin_memory(), "B": Bm.pin_memory()}
        model.data_dependent_initialize(data)
        model.setup(opt)
        model.parallelize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    model.set_input(data)

    def step_resident():
        model.optimize_parameters()

    def step_e2e():
        model.set_input(data)
        model.optimize_parameters()
        model.get_current_losses()

    for _ in range(args.warmup):
        step_resident()
    # the first second after start-up runs ~8 % slower than steady state (allocator growth, clock ramp): keep warming,
    # untimed, for a fixed number of further steps (~1 s) so that the K timed steps below measure the steady state
    for _ in range(EXTRA_WARMUP_2D):            # a fixed count: every rank must run the same number of all-reduces
        step_resident()
    # the whole step (forward, losses, backward, gradient all-reduce, Adam) as one CUDA graph: same kernels, no
    # per-launch host work (DFMIR_CUDA_GRAPH=0: eager launches)
    graphed, graph_note = False, None
    if use_graph:
        try:
            model.capture_step()
            graphed = True
        except Exception as exc:                 # stay on the eager launches of the same kernels
            model._graph = None
            graph_note = f"capture failed, eager launches: {type(exc).__name__}: {str(exc)[:200]}"
            print("bench.py: " + graph_note, file=sys.stderr)
            torch.cuda.synchronize()
        for _ in range(3):
            step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.launch_count_reset()
    ms = timed(step_resident, args.steps)
    launches = model.graph_launches_per_step * args.steps if graphed else _lib.launch_count()
    ms_e2e = timed(step_e2e, args.steps)
    model._graph = None                          # the per-kernel event pass below needs eager launches
    # per-kernel durations for the roofline: the same K steps once more with a CUDA-event pair around every convolution
    # launch (kept out of the timed regions above: ~1000 event records per step cost host time the step no longer hides)
    prof = Fn.ConvProfile()
    Fn.PROFILE = prof
    ms_prof = timed(step_resident, args.steps)
    Fn.PROFILE = None
    kinds = prof.by_kind()
    conv_ms, conv_flops, conv_calls = prof.total()
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        return
    pk, pk_src = peaks()
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    # dominant kernel: the tcgen05 implicit-GEMM convolution (forward and data gradient are the same kernel)
    dom_ms = sum(kinds.get(k, (0, 0, 0))[0] for k in ("umma_fwd", "umma_dgrad"))
    dom_fl = sum(kinds.get(k, (0, 0, 0))[1] for k in ("umma_fwd", "umma_dgrad"))
    dom_n = sum(kinds.get(k, (0, 0, 0))[2] for k in ("umma_fwd", "umma_dgrad"))
    achieved = dom_fl / (dom_ms / 1e3) / 1e12 if dom_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_dominant_kernel.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
    per_kind = {k: {"ms_per_step": v[0] / args.steps, "tflops": (v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 else 0.0),
                    "launches_per_step": v[2] / args.steps} for k, v in sorted(kinds.items())}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if prof.umma_calls else "f32", "data": "synthetic",
        "config": {"workload": f"2D {S}x{S} batch={B}/GPU translation+registration fwd/bwd+Adam (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "conv_engine": Fn.CONV_ENGINE, "arithmetic": "fp32 storage; convolutions with 64..256 channels on tcgen05 "
                   "kind::tf32 (operands truncated to TF32, fp32 accumulate: the class cuDNN runs the reference in); "
                   "every other kernel fp32",
                   "schedule": "feat_k of real_A / real_B tapped from the full generator pass (identical values; the reference "
                               "recomputes them with 3 more encoder passes) - DFMIR_REUSE_REAL_FEATURES=0 restores that",
                   "l2_policy": "inputs larger than L2: each step streams > 10 GB of activations (L2 is 126 MB)",
                   "warmup_note": f"{args.warmup} + {EXTRA_WARMUP_2D} untimed steps before the timed region",
                   "cuda_graph": graphed, **({"cuda_graph_note": graph_note} if graph_note else {})},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(2 * B * S * S * 4), "d2h_bytes_per_step": 6 * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "conv_umma_pair_kernel / conv_umma_halo_kernel / conv_umma_kernel (tcgen05 implicit-GEMM convolution, forward + data gradient)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": f"bf16_tflops_sustained, {pk_src}; kind::tf32 issues at half the bf16 rate, so the "
                                    "TF32 ceiling of this kernel is peak/2",
                     "frac_of_tf32_ceiling": achieved / (peak / 2.0),
                     "flops_per_launch": dom_fl / dom_n if dom_n else 0.0, "launch_ms": dom_ms / dom_n if dom_n else 0.0,
                     "launches_per_step": dom_n / args.steps, "share_of_step": dom_ms / ms if ms > 0 else None,
                     "measured": f"CUDA events around every launch of these kernels over {args.steps} further steps of the same workload "
                                 f"({ms_prof / args.steps:.1f} ms/step with the event records)",
                     "traffic": traffic,
                     "traffic_note": "dram bytes of one ResnetBlock-conv launch (batch 16) from profiles/r1_dominant_kernel.json"
                                     if traffic else None},
        "kernels": per_kind,
        "conv_total": {"ms_per_step": conv_ms / args.steps, "tflops": conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0,
                       "flops_per_step": conv_flops / args.steps, "share_of_step": conv_ms / ms if ms > 0 else None},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(S)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


EXTRA_WARMUP_2D = 12
EXTRA_WARMUP_3D = 40


def run_ours_3d(args):
    """BASELINE configs[2]: VoxelMorph-3D (6-level features) on 128^3 volume pairs, batch 2 per GPU: U-Net fwd,
    fused integrate -> resize -> warp -> NCC + Grad (one cooperative launch), backward, Adam.  Weak scaling."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dfmir_b200 import _lib, vxm
    B = args.batch3d
    shape = tuple(int(v) for v in args.shape3d.split(",")) if args.shape3d else (args.size3d,) * 3
    vox = shape[0] * shape[1] * shape[2]
    six_level = args.features3d == "6level"
    torch.manual_seed(1234)
    feats = [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]] if six_level else None   # None: vxm/networks.py:9-14 defaults
    R = vxm.VxmDense(shape, feats, int_steps=7, bidir=False).cuda()
    params = [p for p in R.parameters() if p.requires_grad]
    flat = None
    if world > 1:
        for t in list(R.parameters()) + list(R.buffers()):
            dist.broadcast(t.data, src=0)
        flat = torch.zeros(sum(p.numel() for p in params), device="cuda")
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p); off += p.numel()
    optim = torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(77 + rank)
    def vol():
        x = torch.randn(B, 1, *shape, generator=g)
        k = torch.ones(1, 1, 5, 5, 5) / 125.0
        for _ in range(2):
            x = torch.nn.functional.conv3d(x, k, padding=2)
        return torch.tanh(3 * x / x.std()).contiguous().pin_memory()
    hA, hB = vol(), vol()
    state = {"A": hA.cuda(), "B": hB.cuda(), "loss": None}

    def step_resident():
        if flat is not None:
            flat.zero_()
        else:
            optim.zero_grad(set_to_none=False) if params[0].grad is not None else None
        y, flow, ncc, grad = R.forward_with_losses(state["A"], state["B"], win=9)
        loss = ncc + 0.02 * grad
        loss.backward()
        if flat is not None:
            dist.all_reduce(flat); flat.mul_(1.0 / world)
        optim.step()
        state["loss"] = loss.detach()

    def step_e2e():
        state["A"] = hA.cuda(non_blocking=True); state["B"] = hB.cuda(non_blocking=True)
        step_resident()
        float(state["loss"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup + EXTRA_WARMUP_3D):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.launch_count_reset()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    if rank != 0:
        return
    # VxmDense-3D 128^3 6-level 284.6 GFLOP per pair (95.0 fwd + 189.6 bwd); default features at
    # 160x192x160 1708.1; both scale with the voxel count
    flops = (284.6e9 * vox / 128 ** 3 if six_level else 1708.1e9 * vox / (160 * 192 * 160)) * B
    name = "x".join(str(v) for v in shape)
    cfg_no = 3 if shape == (160, 192, 160) else 2
    pk, pk_src = peaks()
    line = {
        "metric": METRIC, "value": B * world * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "data": "synthetic",
        "config": {"workload": f"3D {name} batch={B}/GPU VoxelMorph-3D ({'6-level' if six_level else 'default'} features) + VecInt + "
                               f"NCC[9^3] + Grad fwd/bwd+Adam (BASELINE configs[{cfg_no}])",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "arithmetic": "fp32 storage; stride-1 convolutions with >= 16 channels on tcgen05 kind::tf32 (forward, data and weight "
                                 "gradient, 5-D TMA boxes); the stride-2 encoder, the 2-channel first layer and the backward of the planar "
                                 "flow head on fp32 CUDA cores; warp / VecInt / NCC / Grad fp32",
                   "l2_policy": "inputs larger than L2: full-resolution activations are 34 channels x 8 MB per volume"},
        "clocks": clocks,
        "e2e": {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(2 * B * vox * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "all convolution kernels of the step (conv_umma_halo_kernel, conv_wgrad_umma_kernel, fp32 kernels for the strided / thin layers)",
                     "achieved": flops * args.steps / (ms / 1e3) / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": flops * args.steps / (ms / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                     "peak_source": f"bf16_tflops_sustained, {pk_src}; whole-step algorithmic conv FLOPs over step time (not a single kernel)",
                     "traffic": None},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="pairs per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--workload", default="2d", choices=["2d", "3d"],
                    help="2d: BASELINE configs[1] (the headline line, default); 3d: configs[2], VoxelMorph-3D 128^3")
    ap.add_argument("--batch3d", type=int, default=2)
    ap.add_argument("--size3d", type=int, default=128)
    ap.add_argument("--shape3d", default="", help="D,H,W of the 3-D workload (overrides --size3d), e.g. 160,192,160 for configs[3]")
    ap.add_argument("--features3d", default="6level", choices=["6level", "default"],
                    help="VoxelMorph-3D U-Net features: the 6-level list of configs[2] or the reference defaults (configs[3])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the dfmir_b200 path has no CPU fallback (use --impl reference for the CPU arm)")
    if args.workload == "3d":
        run_ours_3d(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
