#!/usr/bin/env python
"""N ranks x batch b  ==  1 rank x batch N*b  ("Distributed"; reference models/base_model.py:103-107: the
reference splits ONE batch over the GPUs with nn.DataParallel, so its gradients are those of the whole batch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp_equiv.py

Every rank (a) takes its slice of a fixed global batch through REGISTRATIONModel.parallelize() + one step (gradients
averaged over NCCL inside the step) and (b) runs the whole global batch through a second, unparallelised model on its
own GPU; both start from the same weights and draw the same patch ids.  Compared after one step on the exact-fp32 engine
(TF32 truncation would hide a wrong scale under its own noise): the six logged losses (rank mean vs global value), every
averaged gradient and the parameters after Adam (where the gradient is resolved above summation noise).  Exit code 0 = equal to summation-order tolerance."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as gi                                    # noqa: E402
from oracle import torch_port as tp                    # noqa: E402  (random state-dicts only: test infrastructure)
import dfmir_b200.functional as Fn                     # noqa: E402
from dfmir_b200 import registration_model as rm        # noqa: E402

S, PER_RANK = 64, 2


class FixedRandperm:
    """Call k of a step returns a fixed permutation (same on every rank, same for both models)."""

    def __init__(self):
        self.k, self.cache = 0, {}

    def __call__(self, n, device=None, generator=None, **kw):
        key = (self.k % 15, int(n))
        self.k += 1
        if key not in self.cache:
            self.cache[key] = torch.from_numpy(np.random.RandomState(4000 + key[0]).permutation(int(n)))
        return self.cache[key].to(device or "cpu")


def build(B, sds, local):
    opt = rm.default_options(batch_size=B, crop_size=S, load_size=S, gpu_ids=[local], ngf=16)
    with contextlib.redirect_stdout(io.StringIO()):
        m = rm.REGISTRATIONModel(opt)
        m.data_dependent_initialize({'A': torch.zeros(B, 1, S, S), 'B': torch.zeros(B, 1, S, S)})
        m.setup(opt)
    for n, sd in zip('GFR', sds):
        getattr(m, 'net' + n).load_state_dict(sd, strict=False)
    return m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Fn.CONV_ENGINE = "simt"
    G = world * PER_RANK
    sds = tp.random_state_dicts(ngf=16, n_blocks=9, crop=S, seed=11)
    sds[2]['flow.weight'] = sds[2]['flow.weight'] * 2e4
    A = torch.from_numpy(gi.image_textured(900, G, (S, S)))
    Bm = torch.from_numpy(gi.image_textured(910, G, (S, S)))
    rp = FixedRandperm()
    real_randperm, torch.randperm = torch.randperm, rp
    try:
        sl = slice(rank * PER_RANK, (rank + 1) * PER_RANK)
        m = build(PER_RANK, sds, local)
        m.parallelize()
        rp.k = 0
        m.set_input({'A': A[sl], 'B': Bm[sl]})
        m.optimize_parameters()
        ref = build(G, sds, local)
        rp.k = 0
        ref.set_input({'A': A, 'B': Bm})
        ref.optimize_parameters()
    finally:
        torch.randperm = real_randperm
    worst = {"loss": 0.0, "grad": 0.0, "param": 0.0}
    la, lb = m.get_current_losses(), ref.get_current_losses()
    for k in lb:
        t = torch.tensor([la[k]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        worst["loss"] = max(worst["loss"], abs(float(t) / world - lb[k]) / max(1.0, abs(lb[k])))
    for name in 'GFR':
        for (k, p), (_, q) in zip(getattr(m, 'net' + name).named_parameters(), getattr(ref, 'net' + name).named_parameters()):
            if p.grad is None or q.grad is None:
                continue
            sc = float(q.grad.abs().max())
            if sc > 1e-20 and not (name == 'G' and k.endswith('.bias')):       # biases in front of an instance norm: true gradient 0
                worst["grad"] = max(worst["grad"], float((p.grad - q.grad).abs().max()) / sc)
            # Adam's first step moves every element by lr * sign(g) whatever |g| is: where the gradient is summation
            # noise around zero the sign, and so the parameter, is arbitrary - compare where |g| is resolved
            live = q.grad.abs() > 1e-3 * sc if sc > 1e-20 else torch.zeros_like(q.grad, dtype=torch.bool)
            if name == 'G' and k.endswith('.bias'):
                live = torch.zeros_like(live)
            if bool(live.any()):
                worst["param"] = max(worst["param"], float((p - q).abs()[live].max()))
    t = torch.tensor([worst["loss"], worst["grad"], worst["param"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t[0]) <= 1e-5 and float(t[1]) <= 2e-3 and float(t[2]) <= 2e-5
    if rank == 0:
        print(f"world {world} x batch {PER_RANK} vs 1 x batch {G} (64x64, ngf 16, 9 blocks, exact-fp32 engine): "
              f"max |d loss| / scale {float(t[0]):.2e}, max |d grad| / max|grad| {float(t[1]):.2e}, "
              f"max |d param| after Adam {float(t[2]):.2e} -> {'EQUAL' if ok else 'DIFFERENT'}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
