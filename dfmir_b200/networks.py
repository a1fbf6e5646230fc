"""Host-side mirror of the reference's CUT networks on the translation hot path
(models/networks.py: define_G :218-268, ResnetGenerator :956-1051, ResnetBlock :1164-1221,
Downsample :37-60, Upsample :73-93, define_F :276-289, PatchSampleF :575-624, Normalize :493-502,
init_net / init_weights :163-215, get_norm_layer :113-132, get_scheduler :135-160).

Same constructors, forward signatures and state-dict keys as the reference.  The module tree below
only HOLDS parameters (nn.Conv2d / buffers at the reference's indices, so checkpoints and
initialisers are interchangeable); forward() never calls those modules.  It runs its own fused
channels-last schedule through libdfmir_b200.so: conv -> (stats, normalise+ReLU written with the
reflected halo the next conv reads) -> conv ..., with no ReflectionPad / InstanceNorm / ReLU
round trips of their own.
"""
import functools

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import _lib
from . import functional as Fn


def get_filter(filt_size=3):
    a = {1: [1.], 2: [1., 1.], 3: [1., 2., 1.], 4: [1., 3., 3., 1.], 5: [1., 4., 6., 4., 1.],
         6: [1., 5., 10., 10., 5., 1.], 7: [1., 6., 15., 20., 15., 6., 1.]}[filt_size]
    a = np.array(a)
    filt = torch.Tensor(a[:, None] * a[None, :])
    return filt / torch.sum(filt)


class Downsample(nn.Module):
    """Anti-aliased blur-pool (reference :37-60); holds the `filt` buffer, computes in blur_down kernels."""

    def __init__(self, channels, pad_type='reflect', filt_size=3, stride=2, pad_off=0):
        super().__init__()
        if filt_size != 3 or stride != 2 or pad_off != 0 or pad_type not in ('refl', 'reflect'):
            raise NotImplementedError("dfmir_b200 Downsample implements the configuration ResnetGenerator uses "
                                      "(reflect pad, 3-tap filter, stride 2)")
        self.channels = channels
        self.register_buffer('filt', get_filter(filt_size)[None, None, :, :].repeat((channels, 1, 1, 1)))

    def forward_cl(self, x):
        return Fn.blur_down_cl(x)

    def forward(self, inp):
        return Fn.blur_down_cl(inp.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class Upsample(nn.Module):
    """Anti-aliased x2 upsampling (reference :73-93)."""

    def __init__(self, channels, pad_type='repl', filt_size=4, stride=2):
        super().__init__()
        if filt_size != 4 or stride != 2 or pad_type not in ('repl', 'replicate'):
            raise NotImplementedError("dfmir_b200 Upsample implements the configuration ResnetGenerator uses "
                                      "(replicate pad, 4-tap filter, stride 2)")
        self.channels = channels
        self.register_buffer('filt', (get_filter(filt_size) * (stride ** 2))[None, None, :, :].repeat((channels, 1, 1, 1)))

    def forward_cl(self, x):
        return Fn.blur_up_cl(x)

    def forward(self, inp):
        return Fn.blur_up_cl(inp.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class Identity(nn.Module):
    def forward(self, x):
        return x


def get_norm_layer(norm_type='instance'):
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == 'none':
        return lambda x: Identity()
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def get_scheduler(optimizer, opt):
    if opt.lr_policy == 'linear':
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    if opt.lr_policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    if opt.lr_policy == 'cosine':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.n_epochs, eta_min=0)
    return NotImplementedError('learning rate policy [%s] is not implemented', opt.lr_policy)


def init_weights(net, init_type='normal', init_gain=0.02, debug=False):
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            if init_type == 'normal':
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm2d') != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


def init_net(net, init_type='normal', init_gain=0.02, gpu_ids=[], debug=False, initialize_weights=True):
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
    if initialize_weights:
        init_weights(net, init_type, init_gain=init_gain, debug=debug)
    return net


class ResnetBlock(nn.Module):
    """Parameter holder with the reference's layout: conv_block = [pad, conv, norm, relu, pad, conv, norm]."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if padding_type != 'reflect' or use_dropout:
            raise NotImplementedError("dfmir_b200 ResnetBlock: reflect padding, no dropout (the reference's defaults)")
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), norm_layer(dim),
            nn.ReLU(True),
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0, bias=use_bias), norm_layer(dim))

    def forward_padded(self, P, out_pad):
        """P: channels-last block input carrying a reflected halo of 1; returns x + conv_block(x)
        written with a halo of out_pad."""
        c1, c2 = self.conv_block[1], self.conv_block[5]
        s1, s2 = Fn.BiasGradSlot(), Fn.BiasGradSlot()      # conv bias gradients summed by the IN backward passes
        rs = Fn.ResidualGradSlot()                         # skip-connection gradient, completed by c1's data gradient
        t1, t2 = Fn.StatsSlot(), Fn.StatsSlot()            # norm statistics summed by the convolution epilogues
        y = Fn.conv_cl(P, c1.weight, c1.bias, bias_slot=s1, res_slot=rs, stats_slot=t1)
        P1 = Fn.instnorm_cl(y, relu=True, out_pad=1, bias_slot=s1, stats_slot=t1)
        y = Fn.conv_cl(P1, c2.weight, c2.bias, bias_slot=s2, stats_slot=t2)
        return Fn.instnorm_cl(y, relu=False, out_pad=out_pad, res=P, res_pad=1, bias_slot=s2, res_slot=rs, stats_slot=t2)

    def forward(self, x):
        P = Fn.pad_reflect_cl(x.permute(0, 2, 3, 1), 1)
        return self.forward_padded(P, 0).permute(0, 3, 1, 2)


class ResnetGenerator(nn.Module):
    """Resnet-based translation generator (reference :956-1051); see module docstring."""

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type='reflect', no_antialias=False, no_antialias_up=False, opt=None):
        assert n_blocks >= 0
        super().__init__()
        self.opt = opt
        if type(norm_layer) == functools.partial:
            use_bias = norm_layer.func == nn.InstanceNorm2d
            is_instance = norm_layer.func == nn.InstanceNorm2d
        else:
            use_bias = norm_layer == nn.InstanceNorm2d
            is_instance = use_bias
        if not is_instance or no_antialias or no_antialias_up or use_dropout or padding_type != 'reflect':
            raise NotImplementedError(
                "dfmir_b200 ResnetGenerator implements the reference's training configuration: "
                "--normG instance, anti-aliased down/up-sampling, reflect padding, no dropout")
        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias),
                 norm_layer(ngf), nn.ReLU(True)]
        n_down = 2
        for i in range(n_down):
            mult = 2 ** i
            model += [nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=1, padding=1, bias=use_bias),
                      norm_layer(ngf * mult * 2), nn.ReLU(True), Downsample(ngf * mult * 2)]
        mult = 2 ** n_down
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer,
                                  use_dropout=use_dropout, use_bias=use_bias)]
        for i in range(n_down):
            mult = 2 ** (n_down - i)
            model += [Upsample(ngf * mult),
                      nn.Conv2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=1, padding=1, bias=use_bias),
                      norm_layer(int(ngf * mult / 2)), nn.ReLU(True)]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*model)
        self.n_blocks, self.n_down = n_blocks, n_down

    @staticmethod
    def _nchw(x_cl):
        return x_cl.permute(0, 3, 1, 2)

    def forward(self, input, layers=[], encode_only=False):
        if -1 in layers:
            layers.append(len(self.model))
        _lib.require_cuda(input)
        m = self.model
        want = set(layers)
        last = layers[-1] if len(layers) > 0 else None
        feats = {}

        class _Stop(Exception):
            pass

        def tap(idx, x_cl, padded=None):
            """record the output of reference layer `idx` (channels-last tensor or view; padded = (P, p) when x_cl is
            the interior of the reflect-padded buffer P)"""
            if idx in want:
                v = self._nchw(x_cl)
                if x_cl._base is None and x_cl.is_contiguous():
                    v._dfmir_cl = x_cl          # lets PatchSampleF's gather return a sparse gradient (functional.gather_patches)
                elif padded is not None and padded[1] > 0 and padded[0].is_contiguous():
                    v._dfmir_cl_pad = padded    # the same for an interior view: sparse gradient on the padded buffer
                feats[idx] = v
            if encode_only and idx == last:
                raise _Stop()

        def interior(P, p):
            return P[:, p:P.shape[1] - p, p:P.shape[2] - p, :] if p else P

        nb, nd = self.n_blocks, self.n_down
        try:
            x = input.permute(0, 2, 3, 1)
            P = Fn.pad_reflect_cl(x, 3); tap(0, P)
            def slot(i):
                """bias-gradient hand-over conv -> IN, unless the conv output is also a feature tap"""
                return None if i in want else Fn.BiasGradSlot()

            sl = slot(1)
            y = Fn.conv_cl(P, m[1].weight, m[1].bias, bias_slot=sl); tap(1, y)
            a = Fn.instnorm_cl(y, relu=True, bias_slot=sl); tap(2, a); tap(3, a)
            idx = 4
            for i in range(nd):
                sl, st = slot(idx), Fn.StatsSlot()
                y = Fn.conv_cl(a, m[idx].weight, m[idx].bias, pad=1, bias_slot=sl, stats_slot=st); tap(idx, y)
                a = Fn.instnorm_cl(y, relu=True, bias_slot=sl, stats_slot=st); tap(idx + 1, a); tap(idx + 2, a)
                a = Fn.blur_down_cl(a); tap(idx + 3, a)
                idx += 4
            if nb > 0:
                P = Fn.pad_reflect_cl(a, 1)
                for b in range(nb):
                    op = 1 if b + 1 < nb else 0
                    P = m[idx].forward_padded(P, op); tap(idx, interior(P, op), padded=(P, op))
                    idx += 1
                a = P
            for i in range(nd):
                a = Fn.blur_up_cl(a); tap(idx, a)
                sl, st = slot(idx + 1), Fn.StatsSlot()
                y = Fn.conv_cl(a, m[idx + 1].weight, m[idx + 1].bias, pad=1, bias_slot=sl, stats_slot=st); tap(idx + 1, y)
                op = 3 if i + 1 == nd else 0
                a = Fn.instnorm_cl(y, relu=True, out_pad=op, bias_slot=sl, stats_slot=st); tap(idx + 2, interior(a, op)); tap(idx + 3, interior(a, op))
                idx += 4
            tap(idx, a)                                     # ReflectionPad2d(3) output
            if (idx + 1) in want:
                y = Fn.conv_cl(a, m[idx + 1].weight, m[idx + 1].bias); tap(idx + 1, y)
                out = torch.tanh(y)
            else:
                out = Fn.conv_cl(a, m[idx + 1].weight, m[idx + 1].bias, act=Fn.ACT_TANH)
            tap(idx + 2, out)
            if len(m) in want:
                feats[len(m)] = self._nchw(out)
        except _Stop:
            return [feats[i] for i in layers if i in feats]
        fake = self._nchw(out)
        if len(layers) > 0:
            return fake, [feats[i] for i in layers if i in feats]
        return fake


def define_G(input_nc, output_nc, ngf, netG, norm='batch', use_dropout=False, init_type='normal', init_gain=0.02,
             no_antialias=False, no_antialias_up=False, gpu_ids=[], opt=None):
    norm_layer = get_norm_layer(norm_type=norm)
    blocks = {'resnet_9blocks': 9, 'resnet_6blocks': 6, 'resnet_4blocks': 4}
    if netG not in blocks:
        raise NotImplementedError('Generator model name [%s] is not on the dfmir_b200 hot path (resnet_{4,6,9}blocks)' % netG)
    net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                          no_antialias=no_antialias, no_antialias_up=no_antialias_up, n_blocks=blocks[netG], opt=opt)
    return init_net(net, init_type, init_gain, gpu_ids, initialize_weights=True)


class Normalize(nn.Module):
    def __init__(self, power=2):
        super().__init__()
        if power != 2:
            raise NotImplementedError("dfmir_b200 Normalize: L2 only")
        self.power = power

    def forward(self, x):
        return Fn.l2norm_rows(x)


class PatchSampleF(nn.Module):
    """Patch sampler + 2-layer MLP + L2 normalisation (reference :575-624)."""

    def __init__(self, use_mlp=False, init_type='normal', init_gain=0.02, nc=256, gpu_ids=[]):
        super().__init__()
        self.l2norm = Normalize(2)
        self.use_mlp = use_mlp
        self.nc = nc
        self.mlp_init = False
        self.init_type = init_type
        self.init_gain = init_gain
        self.gpu_ids = gpu_ids
        self.generator = None       # optional torch.Generator for the patch ids (REGISTRATIONModel.parallelize)

    def create_mlp(self, feats):
        for mlp_id, feat in enumerate(feats):
            input_nc = feat.shape[1]
            mlp = nn.Sequential(*[nn.Linear(input_nc, self.nc), nn.ReLU(), nn.Linear(self.nc, self.nc)])
            mlp.to(feat.device)
            setattr(self, 'mlp_%d' % mlp_id, mlp)
        init_net(self, self.init_type, self.init_gain, self.gpu_ids)
        self.mlp_init = True

    def forward(self, feats, num_patches=64, patch_ids=None):
        return_ids, return_feats = [], []
        if self.use_mlp and not self.mlp_init:
            self.create_mlp(feats)
        for feat_id, feat in enumerate(feats):
            B, C, H, W = feat.shape
            if num_patches <= 0:
                raise NotImplementedError("dfmir_b200 PatchSampleF: num_patches must be > 0 (the reference trains with 256)")
            if patch_ids is not None:
                patch_id = patch_ids[feat_id]
            else:
                patch_id = torch.randperm(H * W, device=feat.device, generator=self.generator) \
                    if self.generator is not None else torch.randperm(H * W, device=feat.device)
                patch_id = patch_id[:int(min(num_patches, patch_id.shape[0]))]
            x_sample = Fn.gather_patches(feat, patch_id)          # (B*P, C)
            if self.use_mlp:
                mlp = getattr(self, 'mlp_%d' % feat_id)
                x_sample = Fn.linear(x_sample, mlp[0].weight, mlp[0].bias, relu=True)
                x_sample = Fn.linear(x_sample, mlp[2].weight, mlp[2].bias)
            return_ids.append(patch_id)
            return_feats.append(self.l2norm(x_sample))
        return return_feats, return_ids


def define_F(input_nc, netF, norm='batch', use_dropout=False, init_type='normal', init_gain=0.02, no_antialias=False,
             gpu_ids=[], opt=None):
    if netF == 'sample':
        net = PatchSampleF(use_mlp=False, init_type=init_type, init_gain=init_gain, gpu_ids=gpu_ids, nc=opt.netF_nc)
    elif netF == 'mlp_sample':
        net = PatchSampleF(use_mlp=True, init_type=init_type, init_gain=init_gain, gpu_ids=gpu_ids, nc=opt.netF_nc)
    else:
        raise NotImplementedError('projection model name [%s] is not on the dfmir_b200 hot path (sample | mlp_sample)' % netF)
    return init_net(net, init_type, init_gain, gpu_ids)
