"""Host-side mirror of the reference's VoxelMorph registration network
(models/voxelmorph/torchvoxelmorph/networks.py: Unet :16-106, VxmDense :1028-1165, ConvBlock
:1506-1521; modelio.py LoadableModel :36-76).  Same constructors, forward signatures, attribute
names and state-dict keys; nn.ConvNd modules only hold the parameters.  forward() runs a
channels-last schedule through libdfmir_b200.so: conv+LeakyReLU fused, nearest-upsample+concat in
one kernel, the flow head written planar for the integrate -> resize -> warp kernels.
"""
import functools
import inspect

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.normal import Normal

from . import _lib
from . import functional as Fn
from . import layers


def default_unet_features():
    return [[16, 32, 32, 32], [32, 32, 32, 32, 32, 16, 16]]


def store_config_args(func):
    """Saves every constructor argument in self.config (reference modelio.py:7-34, on getfullargspec)."""
    spec = inspect.getfullargspec(func)
    attrs, defaults = spec.args, spec.defaults

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        self.config = {}
        if defaults:
            for attr, val in zip(reversed(attrs), reversed(defaults)):
                self.config[attr] = val
        for attr, val in zip(attrs[1:], args):
            self.config[attr] = val
        for attr, val in kwargs.items():
            self.config[attr] = val
        return func(self, *args, **kwargs)
    return wrapper


class LoadableModel(nn.Module):
    def __init__(self, *args, **kwargs):
        if not hasattr(self, 'config'):
            raise RuntimeError('models that inherit from LoadableModel must decorate the constructor with @store_config_args')
        super().__init__(*args, **kwargs)

    def save(self, path):
        sd = self.state_dict().copy()
        for key in [k for k in sd.keys() if k.endswith('.grid')]:
            sd.pop(key)
        torch.save({'config': self.config, 'model_state': sd}, path)

    @classmethod
    def load(cls, path, device):
        checkpoint = torch.load(path, map_location=torch.device(device))
        model = cls(**checkpoint['config'])
        model.load_state_dict(checkpoint['model_state'], strict=False)
        return model


class ConvBlock(nn.Module):
    """conv(k3, stride, pad 1) + LeakyReLU(0.2) (reference :1506-1521), one fused kernel."""

    def __init__(self, ndims, in_channels, out_channels, stride=1):
        super().__init__()
        Conv = getattr(nn, 'Conv%dd' % ndims)
        self.main = Conv(in_channels, out_channels, 3, stride, 1)
        self.activation = nn.LeakyReLU(0.2)
        self.stride = stride

    def forward_cl(self, x):
        return Fn.conv_cl(x, self.main.weight, self.main.bias, stride=self.stride, pad=1, act=Fn.ACT_LEAKY)

    def forward(self, x):
        nd = x.dim() - 2
        perm_in = (0, *range(2, nd + 2), 1)
        perm_out = (0, nd + 1, *range(1, nd + 1))
        return self.forward_cl(x.permute(*perm_in)).permute(*perm_out)


class Unet(nn.Module):
    def __init__(self, inshape, nb_features=None, nb_levels=None, feat_mult=1):
        super().__init__()
        ndims = len(inshape)
        assert ndims in [1, 2, 3], 'ndims should be one of 1, 2, or 3. found: %d' % ndims
        if ndims == 1:
            raise NotImplementedError("dfmir_b200 Unet: 2-D and 3-D volumes")
        if nb_features is None:
            nb_features = default_unet_features()
        if isinstance(nb_features, int):
            if nb_levels is None:
                raise ValueError('must provide unet nb_levels if nb_features is an integer')
            feats = np.round(nb_features * feat_mult ** np.arange(nb_levels)).astype(int)
            self.enc_nf = feats[:-1]
            self.dec_nf = np.flip(feats)
        elif nb_levels is not None:
            raise ValueError('cannot use nb_levels if nb_features is not an integer')
        else:
            self.enc_nf, self.dec_nf = nb_features
        self.upsample = nn.Upsample(scale_factor=2, mode='nearest')
        prev_nf = 2
        self.downarm = nn.ModuleList()
        for nf in self.enc_nf:
            self.downarm.append(ConvBlock(ndims, prev_nf, nf, stride=2))
            prev_nf = nf
        enc_history = list(reversed(self.enc_nf))
        self.uparm = nn.ModuleList()
        for i, nf in enumerate(self.dec_nf[:len(self.enc_nf)]):
            channels = prev_nf + enc_history[i] if i > 0 else prev_nf
            self.uparm.append(ConvBlock(ndims, channels, nf, stride=1))
            prev_nf = nf
        prev_nf += 2
        self.extras = nn.ModuleList()
        for nf in self.dec_nf[len(self.enc_nf):]:
            self.extras.append(ConvBlock(ndims, prev_nf, nf, stride=1))
            prev_nf = nf

    def forward_cl(self, x):
        """x: channels-last (N,*S,2) -> (N,*S,dec_nf[-1])"""
        x_enc = [x]
        for layer in self.downarm:
            x_enc.append(layer.forward_cl(x_enc[-1]))
        x = x_enc.pop()
        for layer in self.uparm:
            x = layer.forward_cl(x)
            x = Fn.upsample_concat_cl(x, x_enc.pop(), pad_channels_to=4)
        for layer in self.extras:
            x = layer.forward_cl(x)
        return x

    def forward(self, x):
        nd = x.dim() - 2
        y = self.forward_cl(x.permute(0, *range(2, nd + 2), 1))
        return y.permute(0, nd + 1, *range(1, nd + 1))


class VxmDense(LoadableModel):
    """VoxelMorph network for nonlinear registration between two images (reference :1028-1165)."""

    @store_config_args
    def __init__(self, inshape, nb_unet_features=None, nb_unet_levels=None, unet_feat_mult=1, int_steps=7,
                 int_downsize=2, bidir=False, use_probs=False):
        super().__init__()
        self.training = True
        ndims = len(inshape)
        assert ndims in [1, 2, 3], 'ndims should be one of 1, 2, or 3. found: %d' % ndims
        self.unet_model = Unet(inshape, nb_features=nb_unet_features, nb_levels=nb_unet_levels, feat_mult=unet_feat_mult)
        Conv = getattr(nn, 'Conv%dd' % ndims)
        self.flow = Conv(self.unet_model.dec_nf[-1], ndims, kernel_size=3, padding=1)
        self.flow.weight = nn.Parameter(Normal(0, 1e-5).sample(self.flow.weight.shape))
        self.flow.bias = nn.Parameter(torch.zeros(self.flow.bias.shape))
        if use_probs:
            raise NotImplementedError('Flow variance has not been implemented in pytorch - set use_probs to False')
        resize = int_steps > 0 and int_downsize > 1
        self.resize = layers.ResizeTransform(int_downsize, ndims) if resize else None
        self.fullsize = layers.ResizeTransform(1 / int_downsize, ndims) if resize else None
        self.bidir = bidir
        down_shape = [int(dim / int_downsize) for dim in inshape]
        self.integrate = layers.VecInt(down_shape, int_steps) if int_steps > 0 else None
        self.transformer = layers.SpatialTransformer(inshape)

    def forward(self, source, target, registration=False):
        _lib.require_cuda(source, target)
        nd = source.dim() - 2
        # channels-last (N,*S,2) input of the U-Net: interleave the two single-channel volumes
        x = torch.stack([source[:, 0], target[:, 0]], dim=-1) if source.shape[1] == 1 and target.shape[1] == 1 \
            else torch.cat([source, target], dim=1).permute(0, *range(2, nd + 2), 1).contiguous()
        x = self.unet_model.forward_cl(x)
        # flow head written planar (N,nd,*S): the layout of the integration / warp kernels
        flow_field = Fn.conv_cl(x, self.flow.weight, self.flow.bias, pad=1, planar_out=True)
        pos_flow = flow_field
        if self.resize:
            pos_flow = self.resize(pos_flow)
        preint_flow = pos_flow
        neg_flow = -pos_flow if self.bidir else None
        if self.integrate:
            if self.bidir:
                # +v and -v are integrated (and resized) by the same launches: virtual batch of 2B
                both = layers.vecint(pos_flow, self.integrate.nsteps, bidir=True)
                if self.fullsize:
                    both = self.fullsize(both)
                B = source.shape[0]
                pos_flow, neg_flow = both[:B], both[B:]
            else:
                pos_flow = self.integrate(pos_flow)
                if self.fullsize:
                    pos_flow = self.fullsize(pos_flow)
        y_source = self.transformer(source, pos_flow)
        y_target = self.transformer(target, neg_flow) if self.bidir else None
        if not registration:
            return (y_source, y_target, pos_flow) if self.bidir else (y_source, preint_flow)
        return y_source, pos_flow

    def _velocity(self, source, target):
        """U-Net + flow head + ResizeTransform(int_downsize): the half-resolution stationary velocity field."""
        nd = source.dim() - 2
        x = torch.stack([source[:, 0], target[:, 0]], dim=-1) if source.shape[1] == 1 and target.shape[1] == 1 \
            else torch.cat([source, target], dim=1).permute(0, *range(2, nd + 2), 1).contiguous()
        x = self.unet_model.forward_cl(x)
        flow_field = Fn.conv_cl(x, self.flow.weight, self.flow.bias, pad=1, planar_out=True)
        return self.resize(flow_field) if self.resize else flow_field

    def forward_with_losses(self, source, target, win=9, eps=1e-5, ncc="sqrt_mean", grad_penalty="l2", grad_mult=1.0):
        """Registration forward with its two losses in ONE launch after the flow head (csrc/fused_reg.cu):
        integrate -> fullsize resize -> warp(source) -> NCC(y_source, target) + Grad(pos_flow).
        Returns (y_source, pos_flow, ncc_loss, grad_loss); values identical to forward() followed by
        NCC_Loss(kernel_var=[win]*nd) / Grad_Loss(dim=nd).  Needs int_steps > 0 and int_downsize == 2."""
        from .fused import integrate_warp_loss
        _lib.require_cuda(source, target)
        if self.integrate is None or self.resize is None or abs(self.resize.factor - 0.5) > 1e-12:
            raise _lib.DfmirError("forward_with_losses: the fused launch covers int_steps > 0 with int_downsize = 2")
        vel = self._velocity(source, target)
        return integrate_warp_loss(vel, source, target, nsteps=self.integrate.nsteps, win=win, eps=eps, ncc=ncc,
                                   grad_penalty=grad_penalty, grad_mult=grad_mult)

    def predict(self, image, flow, svf=True, **kwargs):
        if svf:
            flow = self.integrate(flow)
            if self.fullsize:
                flow = self.fullsize(flow)
        return self.transformer(image, flow, **kwargs)

    def get_flow_field(self, flow_field):
        if self.integrate:
            flow_field = self.integrate(flow_field)
            if self.fullsize:
                flow_field = self.fullsize(flow_field)
        return flow_field
