"""Fused Adam (csrc/adam.cu): drop-in for the torch.optim.Adam instances of the reference's step
(models/registration_model.py:114-115, 135, 168-171) - one launch per optimizer, step counter and learning rate on the
device so that a captured CUDA graph follows the schedulers."""
import ctypes
import struct

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    """Adam (no weight decay, no amsgrad) with the update of every parameter in ONE kernel launch.

    state[p] = {'step', 'exp_avg', 'exp_avg_sq'} as torch.optim.Adam keeps it (state_dicts interchange; 'step' is one
    device tensor shared by the parameters of the optimizer).  `lr` may be a float or a 0-d device tensor (then the
    kernel reads it on the device: schedulers fill it in place and a captured graph follows)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, capturable=True))   # capturable: load_state_dict keeps 'step' on the device
        self._tables = {}        # group index -> (pointer key, tensors table, work table, n_work)

    def _init_state(self, group):
        params = group['params']
        dev = params[0].device
        step = None
        for p in params:
            s = self.state[p].get('step')
            if torch.is_tensor(s) and s.is_cuda and s.dtype == torch.float32 and s.dim() == 0:
                step = s
                break
        if step is None:
            step = torch.zeros((), dtype=torch.float32, device=dev)
            for p in params:             # a state loaded from a torch.optim.Adam checkpoint: adopt its count
                s = self.state[p].get('step')
                if s is not None:
                    step.fill_(float(s))
                    break
        for p in params:
            st = self.state[p]
            if 'exp_avg' not in st:
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['step'] = step
        return step

    def _table(self, gi, group):
        ps = [p for p in group['params'] if p.grad is not None]
        for p in ps:
            _lib.require_cuda(p, p.grad)
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.is_sparse:
                raise _lib.DfmirError("FusedAdam: parameters and gradients must be dense contiguous float32 CUDA tensors")
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr(),
                     p.numel()) for p in ps)
        # two sets of table buffers per group: eager steps and graph captures never share one, so that an eager step taken
        # after a capture (different gradient addresses) cannot rewrite the pinned tables a captured graph re-uploads
        capturing = torch.cuda.is_current_stream_capturing()
        slot = (gi, capturing)
        hit = self._tables.get(slot)
        if hit is not None and hit[0] == key:
            return hit
        # The tables go up through pinned host buffers with asynchronous copies: legal inside a CUDA-graph capture
        # (where the gradients live at the capture pool's addresses, fixed for every replay) - they become memcpy
        # nodes that re-send the same few KB on each replay.  Buffers are sized once for all parameters of the group
        # (no allocation on later calls, none inside a capture after one eager step).
        chunk = int(_lib.lib().dfmir_adam_chunk_elems())
        allp = group['params']
        dev = allp[0].device
        cap_t = 40 * len(allp)
        cap_w = sum((p.numel() + chunk - 1) // chunk for p in allp)
        def alloc():
            return (torch.zeros(cap_t, dtype=torch.uint8).pin_memory(), torch.zeros((max(cap_w, 1), 2), dtype=torch.int32).pin_memory(),
                    torch.zeros(cap_t, dtype=torch.uint8, device=dev), torch.zeros((max(cap_w, 1), 2), dtype=torch.int32, device=dev))
        if hit is None:
            if capturing:
                raise _lib.DfmirError("FusedAdam: run one eager step before capturing (the table buffers are allocated there)")
            bufs = alloc()
            if (gi, True) not in self._tables:       # the capture set, allocated outside any capture
                self._tables[(gi, True)] = (None, None, None, 0, alloc())
        else:
            bufs = hit[4]
            if not capturing:
                torch.cuda.current_stream(dev).synchronize()      # an earlier upload may still be reading the pinned buffers
        h_tab, h_work, t_tab, w_tab = bufs
        rec = b"".join(struct.pack("<QQQQq", *k) for k in key)
        h_tab[:len(rec)] = torch.frombuffer(bytearray(rec), dtype=torch.uint8) if rec else h_tab[:0]
        work = []
        for ti, k in enumerate(key):
            work += [(ti, c) for c in range((k[4] + chunk - 1) // chunk)]
        if work:
            h_work[:len(work)] = torch.tensor(work, dtype=torch.int32)
        t_tab.copy_(h_tab, non_blocking=True)
        w_tab.copy_(h_work, non_blocking=True)
        hit = (key, t_tab, w_tab, len(work), bufs)
        self._tables[slot] = hit
        return hit

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not group['params']:
                continue
            step = self._init_state(group)
            _, t_tab, w_tab, n_work = self._table(gi, group)[:4]
            lr = group['lr']
            lr_dev = lr if torch.is_tensor(lr) and lr.is_cuda else None
            dbl = ctypes.c_double
            _lib.call("dfmir_adam_multi", t_tab, w_tab, n_work, step, lr_dev, dbl(0.0 if lr_dev is not None else float(lr)),
                      dbl(float(group['betas'][0])), dbl(float(group['betas'][1])), dbl(float(group['eps'])))
            # the kernel writes the parameters behind autograd's back: bump their version counters as torch.optim.Adam's
            # in-place ops do (kernel-layout weight copies are cached per version; saved-tensor checks rely on it)
            torch.autograd.graph.increment_version([p for p in group['params'] if p.grad is not None])
        return loss
