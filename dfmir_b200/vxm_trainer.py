"""Training step of the registration network alone — VxmDense + NCC_Loss + Grad_Loss, the 3-D workloads of
BASELINE configs[2..4] (what BASELINE.md section 2 times on the reference: VxmDense.forward
models/voxelmorph/torchvoxelmorph/networks.py:1102-1145, NCC_Loss / Grad_Loss util/losses.py:81-261, backward, Adam).

The reference has no trainer class for this composition (its 3-D runs drive VxmDense and the loss modules from a
script), so the surface follows its BaseModel convention (models/base_model.py): set_input / optimize_parameters /
get_current_losses / parallelize / save_networks, with REGISTRATIONModel's extensions (capture_step: the whole step as
one CUDA graph; one process per GPU with a bucketed NCCL gradient average).  Every tensor op is a kernel of
libdfmir_b200.so: U-Net convolutions (tcgen05), then ONE launch for integrate -> resize -> warp -> NCC + Grad
(csrc/fused_reg.cu) and its backward kernels.
"""
import os

import torch

from . import _lib, vxm


class VxmRegistrationTrainer:
    def __init__(self, inshape, nb_unet_features=None, int_steps=7, win=9, lambda_grad=0.02, lr=2e-4, betas=(0.5, 0.999),
                 device=None, cuda_graph=False, ncc="sqrt_mean", grad_penalty="l2"):
        self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.netR = vxm.VxmDense(tuple(inshape), nb_unet_features, int_steps=int_steps, bidir=False).to(self.device)
        self.win, self.lambda_grad, self.ncc, self.grad_penalty = int(win), float(lambda_grad), ncc, grad_penalty
        self.cuda_graph = bool(cuda_graph)
        lr_ = torch.tensor(float(lr), dtype=torch.float32, device=self.device) if self.cuda_graph else lr
        self.params = [p for p in self.netR.parameters() if p.requires_grad]
        if os.environ.get("DFMIR_FUSED_ADAM", "1") != "0":
            from .optim import FusedAdam
            self.optimizer = FusedAdam(self.params, lr=lr_, betas=betas)        # csrc/adam.cu: one launch
        else:
            self.optimizer = torch.optim.Adam(self.params, lr=lr_, betas=betas, capturable=self.cuda_graph)
        self.loss_names = ['ncc', 'grad']
        self.loss_ncc = self.loss_grad = None
        self._flat_grad, self._world, self._graph = None, 1, None
        self.source = self.target = None
        self.graph_launches_per_step = 0

    # ---- data parallel: one process per GPU, volumes sharded along the batch axis
    def parallelize(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        self._world = dist.get_world_size()
        for t in list(self.netR.parameters()) + list(self.netR.buffers()):
            dist.broadcast(t.data, src=0)
        self._flat_grad = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=self.device)
        off = 0
        for p in self.params:
            p.grad = self._flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()

    def set_input(self, source, target=None):
        if isinstance(source, dict):
            source, target = source['A'], source['B']
        if self._graph is not None:
            self._static[0].copy_(source, non_blocking=True)
            self._static[1].copy_(target, non_blocking=True)
            self.source, self.target = self._static
        else:
            self.source = source.to(self.device, non_blocking=True)
            self.target = target.to(self.device, non_blocking=True)

    def _step(self):
        if self._flat_grad is not None:
            self._flat_grad.zero_()
        else:
            self.optimizer.zero_grad(set_to_none=False)
        self.warped, self.flow, self.loss_ncc, self.loss_grad = self.netR.forward_with_losses(
            self.source, self.target, win=self.win, ncc=self.ncc, grad_penalty=self.grad_penalty)
        loss = self.loss_ncc + self.lambda_grad * self.loss_grad
        loss.backward()
        if self._flat_grad is not None:
            import torch.distributed as dist
            if dist.get_backend() == 'nccl':
                dist.all_reduce(self._flat_grad, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self._flat_grad)
                self._flat_grad.mul_(1.0 / self._world)
        self.optimizer.step()

    def optimize_parameters(self):
        if self._graph is not None:
            self._graph.replay()
            from . import functional as Fn, umma
            Fn._pack_cache.clear()
            umma._kmajor_cache.clear()
            return
        self._step()

    def capture_step(self):
        """The whole step (U-Net, fused cooperative launch, backward, all-reduce, Adam) as one CUDA graph."""
        if not self.cuda_graph:
            raise _lib.DfmirError("capture_step: construct the trainer with cuda_graph=True (capturable Adam)")
        import gc
        from . import functional as Fn
        self._graph = None
        self._static = (self.source.clone(), self.target.clone())
        self.source, self.target = self._static
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        self.warped = self.flow = self.loss_ncc = self.loss_grad = None
        Fn._pack_cache.clear()
        gc.collect()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._step()
        self.graph_launches_per_step = _lib.launch_count() - n0
        self._graph = graph

    def get_current_losses(self):
        return {'ncc': float(self.loss_ncc.detach()), 'grad': float(self.loss_grad.detach())}

    def save_networks(self, path):
        self.netR.save(path)
        torch.save(self.optimizer.state_dict(), os.path.splitext(path)[0] + "_optim.pth")
