#!/usr/bin/env python
"""bench.py — volume-pairs/sec (fwd+bwd+Adam) of the DFMIR translation+registration training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--size S]

The ONE JSON line carries the headline workload (BASELINE configs[1], below) at top level and, under `workloads`, the
3-D registration workloads of configs[2] (128^3, batch 2 / GPU, 6-level U-Net) and configs[3]'s shape (160x192x160,
batch 1 / GPU, the reference's default features), each with its own value / e2e / kernel roofline and (N = 1) CPU leg,
so that the driver's 1/2/4/8-GPU runs record them too.  At N = 1 it also carries `parity` (the step's six losses, NCC
and warp indices against the oracle on the same inputs, computed outside the timed regions) and
`gpu_library_baseline` (the reference's PyTorch path - cuDNN TF32 + ATen - on the same B200, informational).

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


def synthetic_volumes(B, shape, seed):
    g = torch.Generator().manual_seed(seed)

    def vol():
        x = torch.randn(B, 1, *shape, generator=g)
        k = torch.ones(1, 1, 5, 5, 5) / 125.0
        for _ in range(2):
            x = torch.nn.functional.conv3d(x, k, padding=2)
        return torch.tanh(3 * x / x.std()).contiguous()
    return vol(), vol()


FEATS_6LEVEL = [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]
FEATS_DEFAULT = [[16, 32, 32, 32], [32, 32, 32, 32, 32, 16, 16]]      # vxm/networks.py:9-14
WORKLOADS_3D = {          # name -> (shape, batch per GPU, features, BASELINE config index)
    "3d_128": ((128, 128, 128), 2, "6level", 2),
    "3d_160x192x160": ((160, 192, 160), 1, "default", 3),
}


def workload_name_3d(shape, B, features, cfg_no):
    name = "x".join(str(v) for v in shape)
    return (f"3D {name} batch={B}/GPU VoxelMorph-3D ({'6-level' if features == '6level' else 'default'} features) + VecInt + "
            f"NCC[9^3] + Grad fwd/bwd+Adam (BASELINE configs[{cfg_no}])")


def conv_flops_3d(shape, six_level):
    """VxmDense-3D 128^3 6-level 284.6 GFLOP per pair (95.0 fwd + 189.6 bwd); default features at
    160x192x160 1708.1; both scale with the voxel count."""
    vox = shape[0] * shape[1] * shape[2]
    return 284.6e9 * vox / 128 ** 3 if six_level else 1708.1e9 * vox / (160 * 192 * 160)


def cpu_baseline_2d(S, budget_s=20.0):
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
            "sample": f"{n} steps of 1 pair (batch 1) at {S}x{S} after 1 warm-up, oracle/torch_port.Step (torch CPU fp32, {cores} threads)"}


def cpu_baseline_3d(shape, features, max_steps=1):
    """The reference's composition for the 3-D workloads (VxmDense-3D + NCC_Loss[9^3] + Grad_Loss, fwd + bwd + Adam) on
    the host cores: oracle/torch_port.Step3D, one pair per step, no warm-up (a step takes 10 - 40 s)."""
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats = FEATS_6LEVEL if features == "6level" else FEATS_DEFAULT
    st = tp.Step3D(tp.random_state_dict_r3d(feats), feats)
    A, B = synthetic_volumes(1, shape, 77)
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        st.step(A, B)
        n += 1
    dt = time.perf_counter() - t0
    name = "x".join(str(v) for v in shape)
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} step(s) of 1 pair (batch 1) at {name}, no warm-up, oracle/torch_port.Step3D (torch CPU fp32, {cores} threads)"}


def run_reference(args):
    """The reference's CPU PyTorch path (oracle/torch_port: Step for the 2-D headline, Step3D for the 3-D workloads) on all
    host cores, one pair per step."""
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
    for _ in range(min(args.warmup, 2)):
        st.step(A, B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.step(A, B)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} steps of 1 pair (batch 1) at {S}x{S}, torch CPU fp32, {cores} threads"
    workloads = {}
    if not args.no_3d:
        for name, (shape, _, feats, cfg) in WORKLOADS_3D.items():
            cb = cpu_baseline_3d(shape, feats)
            workloads[name] = {"value": cb["value"], "unit": UNIT, "cpu_baseline": cb,
                               "config": {"workload": workload_name_3d(shape, WORKLOADS_3D[name][1], feats, cfg), "sample": "1 pair (batch 1), 1 step"}}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the arm's workload is the headline workload of `bench.py --impl ours`; a step here is a bounded SAMPLE of it
        # (one of the batch's pairs: the step is independent per pair except for the batch mean of the losses)
        "config": {"workload": f"2D {S}x{S} batch={args.batch}/GPU translation+registration fwd/bwd+Adam (BASELINE configs[1])",
                   "batch_per_gpu": args.batch, "global_batch": args.batch * max(args.gpus, 1), "parallelism": f"dp{max(args.gpus, 1)}",
                   "sample": "1 pair (batch 1) per step on the host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "workloads": workloads}))


class Ctx:
    """Process-group plumbing shared by the legs of one bench run."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks (ms)."""
        self.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        self.barrier()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def committed_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from this round's `ncu --set full`
    capture (profiles/<name>; regenerated by tools/ncu_summary.py whenever the kernel changes)."""
    for fn in (name, name.replace("r2_", "r1_")):
        tpath = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(tpath):
            d = json.load(open(tpath))
            return d.get("traffic_bytes_per_launch"), f"profiles/{fn}: {d.get('kernel', '')} {d.get('note', '')}".strip()
    return None, None


def roofline_block(kinds, steps, step_ms, pk, pk_src, kernel_name, traffic_file):
    """Tensor-pipe roofline of the dominant kernel group (tcgen05 implicit-GEMM forward + data gradient): algorithmic
    FLOPs (2*M*N*K per launch) over the CUDA-event duration of those launches, against the measured bf16 peak."""
    dom = [kinds.get(k, (0.0, 0.0, 0, 0.0)) for k in ("umma_fwd", "umma_dgrad")]
    dom_ms, dom_fl, dom_n, dom_by = (sum(v[i] for v in dom) for i in range(4))
    achieved = dom_fl / (dom_ms / 1e3) / 1e12 if dom_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    traffic, tnote = committed_traffic(traffic_file)
    return {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": f"bf16_tflops_sustained, {pk_src}; kind::tf32 issues at half the bf16 rate, so the TF32 ceiling of "
                           "this kernel is peak/2",
            "frac_of_tf32_ceiling": achieved / (peak / 2.0),
            "flops_per_launch": dom_fl / dom_n if dom_n else 0.0, "launch_ms": dom_ms / dom_n if dom_n else 0.0,
            "launches_per_step": dom_n / steps, "share_of_step": dom_ms / step_ms if step_ms > 0 else None,
            "hbm_view": {"algorithmic_bytes_per_launch": dom_by / dom_n if dom_n else 0.0,
                         "achieved_gbs": dom_by / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0, "peak_gbs": pk["hbm_gbs"],
                         "frac": (dom_by / (dom_ms / 1e3) / 1e9 / pk["hbm_gbs"]) if dom_ms > 0 else 0.0,
                         "note": "operand + result bytes (activations in, activations out, weights) of the same launches"},
            "measured": f"CUDA events around every launch of these kernels over {steps} further steps of the same workload",
            "traffic": traffic, "traffic_note": tnote}


def per_kind(kinds, steps):
    return {k: {"ms_per_step": v[0] / steps, "tflops": (v[1] / (v[0] / 1e3) / 1e12 if v[0] > 0 else 0.0),
                "launches_per_step": v[2] / steps} for k, v in sorted(kinds.items())}


def bench_2d(args, ctx):
    from dfmir_b200 import _lib, registration_model as rm
    from dfmir_b200 import functional as Fn
    world, rank, local = ctx.world, ctx.rank, ctx.local
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
    model.set_input(data)

    def step_resident():
        model.optimize_parameters()

    def step_e2e():
        model.set_input(data)
        model.optimize_parameters()
        model.get_current_losses()

    # the first second after start-up runs ~8 % slower than steady state (allocator growth, clock ramp): keep warming,
    # untimed, for a fixed number of further steps so that the K timed steps below measure the steady state
    for _ in range(args.warmup + EXTRA_WARMUP_2D):      # a fixed count: every rank must run the same number of all-reduces
        step_resident()
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
    ms = ctx.timed(step_resident, args.steps)
    launches = model.graph_launches_per_step * args.steps if graphed else _lib.launch_count()
    ms_e2e = ctx.timed(step_e2e, args.steps)
    losses = model.get_current_losses()
    model._graph = None                          # the per-kernel event pass below needs eager launches
    prof = Fn.ConvProfile()
    Fn.PROFILE = prof
    ms_prof = ctx.timed(step_resident, args.steps)
    Fn.PROFILE = None
    kinds = prof.by_kind()
    conv_ms, conv_flops, conv_calls = prof.total()
    clocks = sampler.stop() if sampler else None
    del model
    rm_cleanup()
    if rank != 0:
        return None
    pk, pk_src = peaks()
    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)
    roof = roofline_block(kinds, args.steps, ms, pk, pk_src,
                          "conv_umma_pair_kernel / conv_umma_halo_kernel / conv_umma_kernel (tcgen05 implicit-GEMM convolution, forward + data gradient)",
                          "r2_dominant_kernel.json")
    roof["measured"] += f" ({ms_prof / args.steps:.1f} ms/step with the event records)"
    return {
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
        "losses_last_step": losses,
        "roofline": roof,
        "kernels": per_kind(kinds, args.steps),
        "conv_total": {"ms_per_step": conv_ms / args.steps, "tflops": conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0,
                       "flops_per_step": conv_flops / args.steps, "share_of_step": conv_ms / ms if ms > 0 else None},
    }


def rm_cleanup():
    import gc
    from dfmir_b200 import functional as Fn, umma, losses
    Fn._pack_cache.clear(); umma._kmajor_cache.clear(); Fn._ws_cache.clear(); losses._ws_cache.clear()
    gc.collect()
    torch.cuda.empty_cache()


EXTRA_WARMUP_2D = 12
EXTRA_WARMUP_3D = 30


def bench_3d(args, ctx, shape, B, features, cfg_no, steps=None):
    """VoxelMorph-3D registration step (dfmir_b200.vxm_trainer.VxmRegistrationTrainer): U-Net forward, ONE launch for
    integrate -> resize -> warp -> NCC + Grad, backward, Adam; batch B per GPU, weak scaling."""
    from dfmir_b200 import _lib
    from dfmir_b200 import functional as Fn
    from dfmir_b200.vxm_trainer import VxmRegistrationTrainer
    world, rank, local = ctx.world, ctx.rank, ctx.local
    steps = steps or args.steps
    six_level = features == "6level"
    vox = shape[0] * shape[1] * shape[2]
    use_graph = os.environ.get("DFMIR_CUDA_GRAPH", "1") != "0"
    torch.manual_seed(1234)
    tr = VxmRegistrationTrainer(shape, FEATS_6LEVEL if six_level else None, int_steps=7, win=9, lambda_grad=0.02,
                                device=torch.device("cuda", local), cuda_graph=use_graph)
    tr.parallelize()
    hA, hB = (t.pin_memory() for t in synthetic_volumes(B, shape, 77 + rank))
    tr.set_input(hA, hB)

    def step_resident():
        tr.optimize_parameters()

    def step_e2e():
        tr.set_input(hA, hB)
        tr.optimize_parameters()
        tr.get_current_losses()

    for _ in range(args.warmup + EXTRA_WARMUP_3D):
        step_resident()
    graphed, graph_note = False, None
    if use_graph:
        try:
            tr.capture_step()
            graphed = True
        except Exception as exc:
            tr._graph = None
            graph_note = f"capture failed, eager launches: {type(exc).__name__}: {str(exc)[:200]}"
            print("bench.py (3d): " + graph_note, file=sys.stderr)
            torch.cuda.synchronize()
        for _ in range(3):
            step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.launch_count_reset()
    ms = ctx.timed(step_resident, steps)
    launches = tr.graph_launches_per_step * steps if graphed else _lib.launch_count()
    ms_e2e = ctx.timed(step_e2e, steps)
    losses = tr.get_current_losses()
    tr._graph = None
    prof = Fn.ConvProfile()
    Fn.PROFILE = prof
    ctx.timed(step_resident, steps)
    Fn.PROFILE = None
    kinds = prof.by_kind()
    conv_ms, conv_flops, _ = prof.total()
    clocks = sampler.stop() if sampler else None
    del tr
    rm_cleanup()
    if rank != 0:
        return None
    flops = conv_flops_3d(shape, six_level) * B
    name = "x".join(str(v) for v in shape)
    pk, pk_src = peaks()
    roof = roofline_block(kinds, steps, ms, pk, pk_src,
                          "conv_umma_dmarch_kernel / conv_umma_halo_kernel (tcgen05 implicit-GEMM 3x3x3 convolution: depth-march kernel "
                          "for the 16..64-channel layers, 5-D TMA halo boxes for the others; forward + data gradient)",
                          f"r2_dominant_kernel_{'3d_128' if six_level else '3d_160'}.json")
    return {
        "value": B * world * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "ms_per_step": ms / steps, "scaling": "weak", "dtype": "tf32",
        "config": {"workload": workload_name_3d(shape, B, features, cfg_no),
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": graphed,
                   **({"cuda_graph_note": graph_note} if graph_note else {}),
                   "l2_policy": "inputs larger than L2: full-resolution activations are 16-36 channels x 8-20 MB per volume"},
        "clocks": clocks,
        "e2e": {"value": B * world * steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(2 * B * vox * 4), "d2h_bytes_per_step": 8},
        "gpu_launches": launches, "losses_last_step": losses,
        "roofline": roof,
        "kernels": per_kind(kinds, steps),
        "conv_total": {"ms_per_step": conv_ms / steps, "share_of_step": conv_ms / ms if ms > 0 else None,
                       "algorithmic_tflops_of_step": flops * steps / (ms / 1e3) / 1e12,
                       "frac_of_bf16_peak": flops * steps / (ms / 1e3) / 1e12 / pk["bf16_tflops_sustained"]},
    }


def parity_block(S=256):
    """The benchmarked network (ngf 64, 9 blocks, default engine) on ONE pair against the oracle, same weights / inputs /
    patch ids, outside any timed region: max |d| of the six logged losses vs oracle/torch_port.Step (CPU fp32 with the
    tensor core's TF32 operand truncation emulated), |dNCC| of the NCC kernel vs the C oracle on the step's own
    (regA, real_B), and the number of mismatching int32 warp corner indices vs the C oracle on the step's flow."""
    import numpy as np_
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import inputs as gi
    from oracle import torch_port as tp, c_oracle as orc
    from dfmir_b200 import registration_model as rm, layers, losses
    orc.build()
    sds = tp.random_state_dicts(ngf=64, n_blocks=9, crop=S, seed=5)
    A = torch.from_numpy(gi.image_textured(701, 1, (S, S)))
    Bm = torch.from_numpy(gi.image_textured(711, 1, (S, S)))
    real_randperm = torch.randperm
    try:
        cnt = [100]
        torch.randperm = gi.det_randperm(cnt)
        tp.TF32_EMULATION = "trunc"
        torch.set_num_threads(os.cpu_count() or 1)
        st = tp.Step(*sds, n_blocks=9, batch_size=1, dvf_image=None)
        want = st.step(A, Bm)
        tp.TF32_EMULATION = None
        opt = rm.default_options(batch_size=1, crop_size=S, load_size=S, gpu_ids=[torch.cuda.current_device()])
        with contextlib.redirect_stdout(sys.stderr):
            m = rm.REGISTRATIONModel(opt)
            m.data_dependent_initialize({'A': A, 'B': Bm})
            m.setup(opt)
        for n, sd in zip(('G', 'F', 'R'), sds):
            getattr(m, 'net' + n).load_state_dict(sd, strict=False)
        cnt[0] = 100
        m.set_input({'A': A, 'B': Bm})
        m.optimize_parameters()
        got = m.get_current_losses()
    finally:
        torch.randperm = real_randperm
        tp.TF32_EMULATION = None
    dl = {k: abs(got[k] - want[k]) for k in want}
    # NCC and warp indices on the step's own tensors: CUDA kernels vs the C oracle
    regA, realB = m.regA.detach(), m.real_B.detach()
    ncc_gpu = float(losses.NCC_Loss('cuda', kernel_var=[9, 9])(regA, realB))
    ncc_ref = float(orc.ncc(regA.cpu().numpy(), realB.cpu().numpy())[0])
    flow = m.netR(m.real_A, m.real_B)[2].detach()
    _, idx = layers.warp_indices(m.real_A, flow)
    _, o_idx = orc.warp(m.real_A.cpu().numpy(), flow.cpu().numpy(), return_idx=True)
    mism = int((idx.cpu().numpy() != o_idx).sum())
    del m
    rm_cleanup()
    return {"workload": f"1 pair at {S}x{S}, ngf 64, 9 blocks, tcgen05 engine, eager launches; same state-dicts, inputs and patch ids",
            "comparator": "oracle/torch_port.Step on the CPU (fp32, TF32_EMULATION=trunc: operands truncated as the tensor core does), "
                          "pinned to the reference's own step by tests/test_oracle_nets.py; C oracle for NCC and warp indices",
            "loss_max_abs_diff": max(dl.values()), "loss_abs_diff": dl, "losses": got, "losses_oracle": want,
            "ncc": ncc_gpu, "ncc_oracle": ncc_ref, "ncc_abs_diff": abs(ncc_gpu - ncc_ref), "ncc_tolerance": 1e-4,
            "warp_index_mismatches": mism, "warp_indices_compared": int(o_idx.size)}


def gpu_library_baseline(B, S, steps=3):
    """Informational ('the Blackwell kernel to beat'): the reference's own PyTorch composition
    (oracle/torch_port.Step = F.conv2d / instance_norm / grid_sample / bmm with autograd, Adam) on the SAME B200 through
    cuDNN (allow_tf32, PyTorch's default for convolutions) and ATen kernels, same batch, CUDA-event timed."""
    from oracle import torch_port as tp
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cudnn.allow_tf32 = True
    sds = [{k: v.to(dev) for k, v in sd.items()} for sd in tp.random_state_dicts(crop=S)]
    A, Bm = (t.to(dev) for t in synthetic_pair(B, S, 1234))
    st = tp.Step(*sds, n_blocks=9, batch_size=B, dvf_image=torch.zeros(B, 3, S, S, device=dev))
    for _ in range(2):
        st.step(A, Bm)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        st.step(A, Bm)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    peak = torch.cuda.max_memory_allocated() / 2 ** 30
    del st, sds
    rm_cleanup()
    return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": B, "steps": steps,
            "what": "oracle/torch_port.Step on cuda: PyTorch eager, cuDNN convolutions with allow_tf32=True, ATen grid_sample / "
                    "instance_norm / bmm, torch.optim.Adam; the reference's schedule (feat_k recomputed, graphs built for it)",
            "peak_mem_gib": peak}


def run_ours(args):
    ctx = Ctx()
    line = bench_2d(args, ctx)
    workloads = {}
    if not args.no_3d:
        for name, (shape, B, feats, cfg) in WORKLOADS_3D.items():
            workloads[name] = bench_3d(args, ctx, shape, B, feats, cfg)
    ctx.close()
    if ctx.rank != 0:
        return
    line["workloads"] = workloads
    if ctx.world == 1:
        if not args.no_parity:
            try:
                line["parity"] = parity_block(args.size)
            except Exception as exc:     # the line must still be printed; a missing block reads as unmeasured
                line["parity"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        if not args.no_library_baseline:
            try:
                line["gpu_library_baseline"] = gpu_library_baseline(args.batch, args.size)
            except Exception as exc:
                line["gpu_library_baseline"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
                rm_cleanup()
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_2d(args.size)
            for name, (shape, _, feats, _cfg) in WORKLOADS_3D.items():
                if workloads.get(name) is not None:
                    workloads[name]["cpu_baseline"] = cpu_baseline_3d(shape, feats)
    print(json.dumps(line))


def run_ours_3d(args):
    """One 3-D workload alone (`--workload 3d`): the top-level line is that workload's."""
    ctx = Ctx()
    shape = tuple(int(v) for v in args.shape3d.split(",")) if args.shape3d else (args.size3d,) * 3
    cfg_no = 3 if shape == (160, 192, 160) else 2
    w = bench_3d(args, ctx, shape, args.batch3d, args.features3d, cfg_no)
    ctx.close()
    if ctx.rank != 0:
        return
    line = {"metric": METRIC, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"}
    line.update(w)
    if ctx.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_3d(shape, args.features3d)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="pairs per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-3d", dest="no_3d", action="store_true", help="skip the 3-D workloads of the default line")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true")
    ap.add_argument("--no-library-baseline", dest="no_library_baseline", action="store_true")
    ap.add_argument("--workload", default="2d", choices=["2d", "3d"],
                    help="2d: BASELINE configs[1] headline + the 3-D workloads under `workloads` (default); 3d: one 3-D workload alone")
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
