"""Host-side mirror of the reference's training-step orchestrator
(models/registration_model.py:34-263 REGISTRATIONModel; models/base_model.py BaseModel lifecycle
:70-256) — the hot path named by BASELINE.json.  Same attributes (`netG`, `netF`, `netR`,
`loss_*`, visuals), same method names and step order, so the reference's train.py loop drives it
unchanged; every tensor op of the step runs in libdfmir_b200.so.

Differences from the reference, all at the host level:
  * `feat_k` encoder passes run under no_grad (the reference builds and discards their graph:
    patchnce.py:17 detaches k);
  * `feat_k` for real_A / real_B is tapped from the full generator pass of forward() (same values, the
    encoder is deterministic and InstanceNorm per-sample) instead of being recomputed by three more
    encoder passes (DFMIR_REUSE_REAL_FEATURES=0 restores the reference's schedule);
  * the two masked-L1 terms fuse the mask construction (registration_model.py:160-161) into the
    loss kernel and return a device scalar 0 for an empty mask instead of forcing a host sync
    (`torch.sum(mask) == 0`, :259);
  * `./deform256.jpg` (:148) is decoded once and cached instead of every step; if the file is
    absent a procedural 256x256 grid is used for the `dvf` visual;
  * parallelize() (base_model.py:103-107) keeps one process per GPU: replicated weights, gradients
    in one flat buffer all-reduced once per step over NCCL (no nn.DataParallel).
"""
import argparse
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, layers, losses, networks, vxm
from .patchnce import PatchNCELoss


def default_options(**overrides):
    """The reference's option defaults for this model (options/base_options.py:26-70,
    options/train_options.py:23-40, registration_model.py:39-67 with CUT_mode=CUT)."""
    opt = argparse.Namespace(
        name='experiment_name', gpu_ids=[0], checkpoints_dir='./checkpoints', isTrain=True,
        input_nc=1, output_nc=1, ngf=64, netG='resnet_9blocks', normG='instance', init_type='xavier',
        init_gain=0.02, no_dropout=True, no_antialias=False, no_antialias_up=False,
        batch_size=1, load_size=256, crop_size=256, preprocess='resize_and_crop', direction='AtoB',
        lr=2e-4, beta1=0.5, beta2=0.999, n_epochs=150, n_epochs_decay=150, lr_policy='linear', epoch_count=1,
        lr_decay_iters=50, continue_train=False, epoch='latest', verbose=False, pretrained_name=None,
        CUT_mode='CUT', lambda_GAN=0.0, lambda_NCE=0.25, nce_idt=True, nce_layers='0,4,8,12,16',
        nce_includes_all_negatives_from_minibatch=False, netF='mlp_sample', netF_nc=256, nce_T=0.07,
        num_patches=256, flip_equivariance=False, gan_mode='lsgan', pool_size=0, cuda_graph=False)
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


# feat_k of real_A / real_B taken from the full generator pass instead of three extra encoder passes
# (identical values; DFMIR_REUSE_REAL_FEATURES=0 restores the reference's schedule pass for pass)
REUSE_REAL_FEATURES = os.environ.get("DFMIR_REUSE_REAL_FEATURES", "1") != "0"
# the three optimizers as csrc/adam.cu launches (DFMIR_FUSED_ADAM=0: torch.optim.Adam, capturable in graph mode)
FUSED_ADAM = os.environ.get("DFMIR_FUSED_ADAM", "1") != "0"
# the three PatchNCE terms of a step through netF / the loss kernels in one batch (DFMIR_BATCH_NCE=0: one call per term)
BATCH_NCE_TERMS = os.environ.get("DFMIR_BATCH_NCE", "1") != "0"

_test_image_cache = {}


def open_image_to_torch(path, size):
    """CenterCrop(size) + ToTensor + Normalize(0.5, 0.5) of an image file (reference :14-23), cached."""
    key = (os.path.abspath(path), size)
    if key not in _test_image_cache:
        if os.path.exists(path):
            from PIL import Image
            img = Image.open(path)
            w, h = img.size
            left, top = int(round((w - size) / 2.0)), int(round((h - size) / 2.0))
            img = img.crop((left, top, left + size, top + size))
            a = np.asarray(img, dtype=np.float32) / 255.0
            if a.ndim == 2:
                a = a[:, :, None]
            t = torch.from_numpy(a).permute(2, 0, 1)
        else:  # procedural deformation grid: lines every 16 pixels, 3 channels
            g = np.ones((size, size), np.float32)
            g[::16, :] = 0.0
            g[:, ::16] = 0.0
            t = torch.from_numpy(np.stack([g, g, g]))
        _test_image_cache[key] = ((t - 0.5) / 0.5).unsqueeze(0).contiguous()
    return _test_image_cache[key]


def global_mask_scale(msum, world):
    """Factor that turns a rank-local masked mean  S_r / M_r  into this rank's share of the global-batch
    masked mean  sum_r S_r / sum_r M_r  (reference registration_model.py:262-263 divides by the mask sum of
    the WHOLE batch): after gradients are averaged over the `world` ranks,
    mean_r[(S_r / M_r) * (M_r * world / sum M)] = sum S / sum M.  `msum` may hold several mask sums (one
    entry per masked term of the step): ONE small all-reduce serves them all."""
    import torch.distributed as dist
    total = msum.clone()
    dist.all_reduce(total)
    return torch.where(total > 0, msum * world / total.clamp_min(1e-20), torch.ones_like(total))


def smooothing_loss(y_pred):
    return losses.smooothing_loss(y_pred)


class BaseModel:
    """The slice of models/base_model.py the training loop touches."""

    def __init__(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        self.device = torch.device('cuda:{}'.format(self.gpu_ids[0])) if self.gpu_ids else torch.device('cpu')
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.visual_names, self.optimizers, self.image_paths = [], [], [], [], []
        self.metric = 0
        self._flat_grad = None
        self._buckets, self._bucket_work, self._overlap = [], [], False
        self._world = 1
        self._graph = None
        self.graph_launches_per_step = 0

    def setup(self, opt):
        if self.isTrain:
            self.schedulers = [networks.get_scheduler(optimizer, opt) for optimizer in self.optimizers]
        if not self.isTrain or opt.continue_train:
            self.load_networks(opt.epoch)
        self.print_networks(opt.verbose)

    def parallelize(self):
        """One process per GPU (torchrun): replicate weights from rank 0 once, then average gradients over NCCL.
        Every `.grad` is a view of one flat buffer laid out in the order the backward pass finishes the
        parameters: [R, F | G decoder half | G encoder half].  Each bucket's all-reduce (ReduceOp.AVG) is
        issued from a post-accumulate hook as soon as its last gradient is written, so the collectives of
        the first buckets overlap the rest of the backward pass; optimize_parameters() waits for them
        before the Adam steps.  The global RNG is left alone (each rank keeps its own data order):
        the patch ids PatchSampleF draws come from a dedicated generator seeded identically on all ranks
        (the reference draws one permutation per layer for the whole DataParallel batch, networks.py:609)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        self._world = dist.get_world_size()
        for name in self.model_names:
            net = getattr(self, 'net' + name)
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t.data, src=0)
        groups = self._grad_bucket_groups()
        params = [p for g in groups for p in g]
        total = sum(p.numel() for p in params)
        dev = params[0].device
        self._flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self._buckets, self._bucket_work, self._bucket_hooks = [], [], []
        off = 0
        for g in groups:
            start = off
            for p in g:
                p.grad = self._flat_grad[off:off + p.numel()].view_as(p)
                off += p.numel()
            if off > start:
                self._buckets.append([self._flat_grad[start:off], len(g), 0, False])   # view, #params, #ready, reduced
        overlap = os.environ.get("DFMIR_OVERLAP_ALLREDUCE", "1") != "0"
        if overlap:
            for bi, g in enumerate(groups):
                for p in g:
                    self._bucket_hooks.append(p.register_post_accumulate_grad_hook(self._make_bucket_hook(bi)))
        self._overlap = overlap
        seed = torch.randint(0, 2 ** 31 - 1, (1,), device=dev)
        dist.broadcast(seed, src=0)
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed.item()))
        netF = getattr(self, 'netF', None)
        if netF is not None:
            netF.generator = gen

    def _grad_bucket_groups(self):
        """Parameters grouped by when the backward pass completes their gradients (see parallelize)."""
        def trainable(net):
            return [p for p in net.parameters() if p.requires_grad]
        early, dec, enc = [], [], []
        for name in self.model_names:
            net = getattr(self, 'net' + name)
            if name != 'G' or not hasattr(net, 'model'):
                early += trainable(net)
                continue
            # modules up to the last feature tap are shared by the encoder passes: their packed-weight
            # gradient is complete only when the full pass has been walked back to them
            last_tap = max(getattr(self, 'nce_layers', [0]))
            for i, m in enumerate(net.model):
                (enc if i <= last_tap else dec).extend(trainable(m))
        return [g for g in (early, dec, enc) if g]

    def _make_bucket_hook(self, bi):
        def hook(_param):
            b = self._buckets[bi]
            b[2] += 1
            if b[2] == b[1] and not b[3]:
                b[3] = True
                self._bucket_work.append(self._allreduce_mean(b[0], True))
        return hook

    def _allreduce_mean(self, t, async_op):
        """Mean over ranks: NCCL averages inside the collective; other backends (gloo in the CPU tests) sum and
        _sync_grads scales afterwards."""
        import torch.distributed as dist
        native = dist.get_backend() == 'nccl'
        self._scale_after = not native
        return dist.all_reduce(t, op=dist.ReduceOp.AVG if native else dist.ReduceOp.SUM, async_op=async_op)

    def _zero_grads(self):
        if self._flat_grad is not None:
            self._flat_grad.zero_()
            for b in self._buckets:
                b[2], b[3] = 0, False
        else:
            for opt_ in self.optimizers:
                opt_.zero_grad()

    def _sync_grads(self):
        if self._flat_grad is None:
            return
        if self._overlap:
            for b in self._buckets:
                if not b[3]:            # a parameter of this bucket received no gradient during the backward pass
                    self._bucket_work.append(self._allreduce_mean(b[0], True))
                b[2], b[3] = 0, False
            for w in self._bucket_work:
                w.wait()
            self._bucket_work = []
        else:
            self._allreduce_mean(self._flat_grad, False)
        if getattr(self, '_scale_after', False):
            self._flat_grad.mul_(1.0 / self._world)

    def data_dependent_initialize(self, data):
        pass

    def eval(self):
        for name in self.model_names:
            getattr(self, 'net' + name).eval()

    def test(self):
        with torch.no_grad():
            self.forward()
            self.compute_visuals()

    def compute_visuals(self):
        pass

    def get_image_paths(self):
        return self.image_paths

    def update_learning_rate(self):
        # with opt.cuda_graph the learning rate is a device tensor the captured Adam kernels read: the
        # schedulers fill it in place and the graph stays valid; a python-float rate is baked into a capture
        if self._graph is not None and not all(torch.is_tensor(g['lr']) for o in self.optimizers for g in o.param_groups):
            print('dfmir_b200: learning rate is not a device tensor; dropping the captured step (eager launches from here on)')
            self._graph = None
        for scheduler in self.schedulers:
            if self.opt.lr_policy == 'plateau':
                scheduler.step(self.metric)
            else:
                scheduler.step()
        print('learning rate = %.7f' % float(self.optimizers[0].param_groups[0]['lr']))

    def get_current_visuals(self):
        return OrderedDict((name, getattr(self, name)) for name in self.visual_names)

    def get_current_losses(self):
        return OrderedDict((name, float(getattr(self, 'loss_' + name).detach() if torch.is_tensor(getattr(self, 'loss_' + name)) else getattr(self, 'loss_' + name))) for name in self.loss_names)

    def save_networks(self, epoch):
        """`<epoch>_net_<name>.pth` per network with the reference's state-dict keys (base_model.py:164-180), plus
        `<epoch>_optim.pth` with the optimiser / scheduler state the reference omits, so that a resumed
        run continues Adam's moments and the learning-rate schedule instead of restarting them."""
        os.makedirs(self.save_dir, exist_ok=True)
        for name in self.model_names:
            net = getattr(self, 'net' + name)
            sd = OrderedDict((k, v.detach().cpu()) for k, v in net.state_dict().items())
            torch.save(sd, os.path.join(self.save_dir, '%s_net_%s.pth' % (epoch, name)))
        if self.isTrain and self.optimizers:
            torch.save({'optimizers': [o.state_dict() for o in self.optimizers],
                        'schedulers': [s.state_dict() for s in getattr(self, 'schedulers', [])]},
                       os.path.join(self.save_dir, '%s_optim.pth' % epoch))

    def load_networks(self, epoch):
        load_dir = os.path.join(self.opt.checkpoints_dir, self.opt.pretrained_name) \
            if self.opt.isTrain and self.opt.pretrained_name is not None else self.save_dir
        for name in self.model_names:
            load_path = os.path.join(load_dir, '%s_net_%s.pth' % (epoch, name))
            print('loading the model from %s' % load_path)
            state_dict = torch.load(load_path, map_location=str(self.device))
            if hasattr(state_dict, '_metadata'):
                del state_dict._metadata
            getattr(self, 'net' + name).load_state_dict(state_dict)
        optim_path = os.path.join(load_dir, '%s_optim.pth' % epoch)
        if self.isTrain and os.path.exists(optim_path):        # absent in checkpoints written by the reference
            st = torch.load(optim_path, map_location=str(self.device))
            for o, sd in zip(self.optimizers, st['optimizers']):
                lrs = [g['lr'] for g in o.param_groups]
                o.load_state_dict(sd)
                for g, lr in zip(o.param_groups, lrs):         # keep the device-tensor learning rate of a captured step
                    if torch.is_tensor(lr):
                        lr.fill_(float(g['lr']))
                        g['lr'] = lr
            for s, sd in zip(getattr(self, 'schedulers', []), st['schedulers']):
                s.load_state_dict(sd)

    def print_networks(self, verbose):
        print('---------- Networks initialized -------------')
        for name in self.model_names:
            net = getattr(self, 'net' + name)
            num_params = sum(p.numel() for p in net.parameters())
            if verbose:
                print(net)
            print('[Network %s] Total number of parameters : %.3f M' % (name, num_params / 1e6))
        print('-----------------------------------------------')


class REGISTRATIONModel(BaseModel):
    def __init__(self, opt):
        BaseModel.__init__(self, opt)
        self.loss_names = ['G', 'NCE', 'R', 'smooth', 'local']
        self.visual_names = ['real_A', 'fake_B', 'real_B', 'dvf', 'registered', 'regA']
        self.nce_layers = [int(i) for i in self.opt.nce_layers.split(',')]
        if opt.nce_idt and self.isTrain:
            self.loss_names += ['NCE_Y']
            self.visual_names += ['idt_B']
        self.model_names = ['G', 'F', 'R'] if self.isTrain else ['G', 'R']
        if opt.lambda_GAN > 0.0:
            raise NotImplementedError("dfmir_b200: the discriminator-free model (lambda_GAN = 0, the reference default)")
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.normG, not opt.no_dropout,
                                      opt.init_type, opt.init_gain, opt.no_antialias, opt.no_antialias_up,
                                      self.gpu_ids, opt)
        self.netF = networks.define_F(opt.input_nc, opt.netF, opt.normG, not opt.no_dropout, opt.init_type,
                                      opt.init_gain, opt.no_antialias, self.gpu_ids, opt)
        nb_features = [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]
        vol_shape = (opt.crop_size, opt.crop_size)
        self.netR = vxm.VxmDense(vol_shape, nb_features, int_steps=7, bidir=True).to(self.device)
        self.netR.train()
        self.spatialTransformer = layers.SpatialTransformer(vol_shape).to(self.device)
        if self.isTrain:
            self.criterionNCE = [PatchNCELoss(opt).to(self.device) for _ in self.nce_layers]
            self.criterionNCC = losses.NCC_Loss(self.device, name='ncc', kernel_var=[9, 9], kernel_type='mean')
            self.optimizer_G = self._make_adam(self.netG.parameters())
            self.optimizer_R = self._make_adam(self.netR.parameters())
            self.optimizers.append(self.optimizer_G)
            self.optimizers.append(self.optimizer_R)

    def _make_adam(self, params):
        """torch.optim.Adam with the reference's hyper-parameters (registration_model.py:114-115,135).  With
        opt.cuda_graph the step counters AND the learning rate live on the device (capturable Adam, tensor lr): a
        captured step keeps following the schedulers, which fill the tensor in place."""
        cap = bool(getattr(self.opt, 'cuda_graph', False))
        lr = torch.tensor(float(self.opt.lr), dtype=torch.float32, device=self.device) if cap else self.opt.lr
        if FUSED_ADAM:       # csrc/adam.cu: one launch per optimizer (torch's capturable Adam: ~6 multi-tensor launches)
            from .optim import FusedAdam
            return FusedAdam(params, lr=lr, betas=(self.opt.beta1, self.opt.beta2))
        return torch.optim.Adam(params, lr=lr, betas=(self.opt.beta1, self.opt.beta2), capturable=cap)

    def data_dependent_initialize(self, data):
        self.set_input(data)
        bs_per_gpu = self.real_A.size(0) // max(len(self.opt.gpu_ids), 1)
        self.real_A = self.real_A[:bs_per_gpu]
        self.real_B = self.real_B[:bs_per_gpu]
        self.forward()
        if self.opt.isTrain:
            self.compute_G_loss().backward()
            if self.opt.lambda_NCE > 0.0:
                self.optimizer_F = self._make_adam(self.netF.parameters())
                self.optimizers.append(self.optimizer_F)

    def optimize_parameters(self):
        if self._graph is not None:
            self._graph.replay()
            # the captured Adam kernels update the parameters in place without bumping their version counters:
            # kernel-layout weight copies cached by a later no-grad forward (test(), visuals) must not survive
            from . import functional as Fn, umma
            Fn._pack_cache.clear()
            umma._kmajor_cache.clear()
            return
        self._step()

    def capture_step(self):
        """Capture one training step (forward, losses, backward, gradient all-reduce, the three Adam steps) into a CUDA
        graph; optimize_parameters() then replays it: ~2.3 k kernel launches per step stop costing host time
        (SURVEY 8f N1).  Needs opt.cuda_graph=True at construction (capturable Adam with the learning rate held in a
        device tensor, so update_learning_rate() keeps the graph), a few eager steps on inputs of the final shape first
        (lazy initialisation, allocator warm-up), and set_input() afterwards copies into the captured input buffers."""
        if not getattr(self.opt, 'cuda_graph', False):
            raise _lib.DfmirError("capture_step: construct the model with opt.cuda_graph=True (capturable Adam)")
        self._graph = None
        self._static_A, self._static_B = self.real_A.clone(), self.real_B.clone()
        self.real_A, self.real_B = self._static_A, self._static_B
        # AccumulateGrad nodes made by earlier steps live on the default stream for as long as any tensor of those
        # steps' tapes is alive, and would tie the capture to the legacy stream: drop the old tapes, warm up on a side
        # stream (torch.cuda.graphs recipe), drop again, then capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self._release_tapes()
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        self._release_tapes()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        gen = getattr(getattr(self, 'netF', None), 'generator', None)
        if gen is not None:             # patch ids drawn from the rank-shared generator advance with every replay
            graph.register_generator_state(gen)
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._step()
        self.graph_launches_per_step = _lib.launch_count() - n0
        self._graph = graph

    def _release_tapes(self):
        import gc
        from . import functional as Fn
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v) and v.grad_fn is not None:
                setattr(self, k, v.detach())
        Fn._pack_cache.clear()
        gc.collect()

    def _step(self):
        self.forward()
        y_output = self.netR(self.real_A, self.real_B)
        pos_flow = y_output[2]
        self.registered = self.spatialTransformer(self.fake_B, pos_flow)
        self.regA = y_output[0]
        test_image = getattr(self, '_test_image_dev', None)
        if test_image is None:      # decoded and uploaded once (the reference re-reads the file every step, :148)
            test_image = self._test_image_dev = open_image_to_torch("./deform256.jpg", self.opt.crop_size).to(self.device)
        with torch.no_grad():
            self.dvf = self.spatialTransformer(test_image.expand(pos_flow.shape[0], -1, -1, -1).contiguous()
                                               if pos_flow.shape[0] > 1 else test_image, pos_flow)
        self._zero_grads()

        if (BATCH_NCE_TERMS and self.opt.lambda_NCE > 0.0 and self.opt.nce_idt and self.netF.use_mlp and self.netF.mlp_init
                and not self.opt.flip_equivariance and not self.opt.nce_includes_all_negatives_from_minibatch):
            # the three PatchNCE terms of the step (:152-158, :218-226) share netF: their sampled rows go through its MLPs
            # and the loss kernels together (same values per row, a third of the launches)
            self.loss_NCE, self.loss_NCE_Y, local = self._nce_losses_batched(
                [(self.real_A, self.fake_B), (self.real_B, self.idt_B), (self.real_B, self.regA)])
            self.loss_G_GAN = 0.0
            self.loss_G = self.loss_G_GAN + (self.loss_NCE + self.loss_NCE_Y) * 0.5
            self.loss_local = local * 0.25
        else:
            self.loss_G = self.compute_G_loss()
            self.loss_local = self.calculate_NCE_loss(self.real_B, self.regA) * 0.25
        # masked L1 terms: mask = (u > -0.95) | (v > -0.95) built inside the loss kernel (:160-161)
        l1_a, l1_b = self._masked_l1_pair((self.registered, self.real_B, self.real_B, self.registered),
                                          (self.idt_B, self.registered, self.idt_B, self.registered))
        self.loss_R = l1_a * 1.0 + l1_b * 1.0 + self.loss_local * 1.0
        self.loss_smooth = smooothing_loss(pos_flow) * 0.20
        all_G_loss = self.loss_R + self.loss_G + self.loss_smooth
        all_G_loss.backward()
        self._sync_grads()
        self.optimizer_G.step()
        self.optimizer_R.step()
        if self.opt.netF == 'mlp_sample':
            self.optimizer_F.step()

    def _masked_l1_pair(self, *terms):
        """The step's masked-L1 terms; across ranks their mask sums are made global by ONE all-reduce."""
        out = [losses.l1_threshold_masked(src, tgt, mu, mv, thr=-0.95, return_mask_sum=True) for src, tgt, mu, mv in terms]
        if self._world > 1:
            scale = global_mask_scale(torch.stack([m.detach() for _, m in out]), self._world)
            return [loss * scale[i] for i, (loss, _) in enumerate(out)]
        return [loss for loss, _ in out]

    def set_input(self, input):
        AtoB = self.opt.direction == 'AtoB'
        if self._graph is not None:         # the captured step reads these buffers
            self._static_A.copy_(input['A' if AtoB else 'B'], non_blocking=True)
            self._static_B.copy_(input['B' if AtoB else 'A'], non_blocking=True)
            self.real_A, self.real_B = self._static_A, self._static_B
        else:
            self.real_A = input['A' if AtoB else 'B'].to(self.device, non_blocking=True)
            self.real_B = input['B' if AtoB else 'A'].to(self.device, non_blocking=True)
        self.image_paths = input.get('A_paths' if AtoB else 'B_paths', [])

    def forward(self):
        self.real = torch.cat((self.real_A, self.real_B), dim=0)
        if self.opt.flip_equivariance:
            self.flipped_for_equivariance = self.opt.isTrain and (np.random.random() < 0.5)
            if self.flipped_for_equivariance:
                self.real = torch.flip(self.real, [3])
        # The full pass over cat(real_A, real_B) already computes the encoder activations the reference
        # recomputes three times as feat_k = netG(real_A / real_B, nce_layers, encode_only=True)
        # (registration_model.py:244; InstanceNorm is per-sample, so the values are the same): tap them
        # here and reuse them, detached, as PatchNCELoss detaches k anyway (patchnce.py:17).
        self._real_feats = None
        flipped = self.opt.flip_equivariance and getattr(self, 'flipped_for_equivariance', False)
        if self.isTrain and REUSE_REAL_FEATURES and not flipped:
            self.fake, feats = self.netG(self.real, list(self.nce_layers))
            self._real_feats = [f.detach() for f in feats]
        else:
            self.fake = self.netG(self.real)
        self.fake_B = self.fake[:self.real_A.size(0)]
        self.idt_B = self.fake[self.real_A.size(0):]

    def compute_G_loss(self):
        if self.opt.lambda_NCE > 0.0:
            self.loss_NCE = self.calculate_NCE_loss(self.real_A, self.fake_B)
        else:
            self.loss_NCE, self.loss_NCE_bd = 0.0, 0.0
        if self.opt.nce_idt and self.opt.lambda_NCE > 0.0:
            self.loss_NCE_Y = self.calculate_NCE_loss(self.real_B, self.idt_B)
            loss_NCE_both = (self.loss_NCE + self.loss_NCE_Y) * 0.5
        else:
            loss_NCE_both = self.loss_NCE
        self.loss_G_GAN = 0.0
        self.loss_G = self.loss_G_GAN + loss_NCE_both
        return self.loss_G

    def calculate_NCE_loss(self, src, tgt, patch_ids=None):
        n_layers = len(self.nce_layers)
        feat_q = self.netG(tgt, self.nce_layers, encode_only=True)
        if self.opt.flip_equivariance and self.flipped_for_equivariance:
            feat_q = [torch.flip(fq, [3]) for fq in feat_q]
        mlp_ready = (not self.netF.use_mlp) or self.netF.mlp_init
        with torch.no_grad():   # k is detached inside PatchNCELoss (patchnce.py:17)
            B = self.real_A.size(0)
            if self._real_feats is not None and src is self.real_A:
                feat_k = [f[:B] for f in self._real_feats]
            elif self._real_feats is not None and src is self.real_B:
                feat_k = [f[B:] for f in self._real_feats]
            else:
                feat_k = self.netG(src, self.nce_layers, encode_only=True)
            if not mlp_ready:
                self.netF.create_mlp(feat_k)
            feat_k_pool, sample_ids = self.netF(feat_k, self.opt.num_patches, patch_ids)
        self._last_patch_ids = sample_ids           # observable by tests (graph replays must draw fresh ids)
        feat_q_pool, _ = self.netF(feat_q, self.opt.num_patches, sample_ids)
        total_nce_loss = 0.0
        for f_q, f_k, crit, nce_layer in zip(feat_q_pool, feat_k_pool, self.criterionNCE, self.nce_layers):
            loss = crit(f_q, f_k) * self.opt.lambda_NCE
            total_nce_loss += loss.mean()
        return total_nce_loss / n_layers

    def _nce_losses_batched(self, pairs):
        """calculate_NCE_loss for several (src, tgt) pairs at once.  Patch ids are drawn in the reference's order (pair by
        pair, layer by layer); per layer the sampled rows of all pairs are concatenated, so that each MLP of netF and the
        PatchNCE kernels run once on len(pairs) * B images (negatives stay per image: the batch count of the loss kernel
        is len(pairs) * B).  Every row sees exactly the arithmetic of the one-pair path."""
        from . import functional as Fn
        n_layers, T = len(self.nce_layers), len(pairs)
        B = self.real_A.size(0)
        P = self.opt.num_patches
        feats_q = [self.netG(tgt, self.nce_layers, encode_only=True) for _, tgt in pairs]
        feats_k = []
        with torch.no_grad():
            for src, _ in pairs:
                if self._real_feats is not None and src is self.real_A:
                    feats_k.append([f[:B] for f in self._real_feats])
                elif self._real_feats is not None and src is self.real_B:
                    feats_k.append([f[B:] for f in self._real_feats])
                else:
                    feats_k.append(self.netG(src, self.nce_layers, encode_only=True))
        ids = []
        for t in range(T):
            row = []
            for feat in feats_k[t]:
                hw = feat.shape[2] * feat.shape[3]
                pid = torch.randperm(hw, device=feat.device, generator=self.netF.generator) if self.netF.generator is not None \
                    else torch.randperm(hw, device=feat.device)
                row.append(pid[:int(min(P, pid.shape[0]))])
            ids.append(row)
        self._last_patch_ids = ids[-1]
        batch = T * B
        totals = [0.0] * T
        for li in range(n_layers):
            mlp = getattr(self.netF, 'mlp_%d' % li)

            def project(rows):
                rows = Fn.linear(rows, mlp[0].weight, mlp[0].bias, relu=True)
                return Fn.l2norm_rows(Fn.linear(rows, mlp[2].weight, mlp[2].bias))
            with torch.no_grad():
                k_pool = project(torch.cat([Fn.gather_patches(feats_k[t][li], ids[t][li]) for t in range(T)], 0))
            q_pool = project(torch.cat([Fn.gather_patches(feats_q[t][li], ids[t][li]) for t in range(T)], 0))
            per_row = Fn.patchnce(q_pool, k_pool, batch, self.opt.nce_T) * self.opt.lambda_NCE
            n = per_row.shape[0] // T
            for t in range(T):
                totals[t] = totals[t] + per_row[t * n:(t + 1) * n].mean()
        return [x / n_layers for x in totals]

    def calculate_L1_loss(self, src, tgt, mask):
        return losses.calculate_L1_loss(src, tgt, mask)


RegistrationModel = REGISTRATIONModel
