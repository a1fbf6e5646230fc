"""CPU restatement of the reference's two networks and its patch sampler, composed from the C
oracle's layers (TEST INFRASTRUCTURE ONLY).  Parameters arrive as a dict of numpy arrays keyed like
the reference's state_dict.

    resnet_generator   models/networks.py:956-1051 (ResnetGenerator.forward incl. layers= taps)
    unet / vxm_dense   models/voxelmorph/torchvoxelmorph/networks.py:88-106, 1102-1145
    patch_sample       models/networks.py:597-624 (given patch ids)
"""
import numpy as np

from . import c_oracle as orc


def relu(x):
    return np.maximum(x, np.float32(0))


def leaky(x):
    return np.where(x > 0, x, np.float32(0.2) * x).astype(np.float32)


def resnet_generator(x, sd, n_blocks, layers=()):
    """Returns (fake, {layer_id: feature}) following the reference's Sequential indices."""
    feats = {}

    def tap(i, v):
        if i in layers:
            feats[i] = v
        return v

    a = tap(0, orc.pad_reflect(x, 3))
    a = tap(1, orc.conv(a, sd['model.1.weight'], sd['model.1.bias']))
    a = relu(orc.instnorm(a)); tap(2, a); tap(3, a)
    idx = 4
    for _ in range(2):
        a = tap(idx, orc.conv(a, sd[f'model.{idx}.weight'], sd[f'model.{idx}.bias'], pad=1))
        a = relu(orc.instnorm(a)); tap(idx + 1, a); tap(idx + 2, a)
        a = tap(idx + 3, orc.blur_down(a))
        idx += 4
    for _ in range(n_blocks):
        p = f'model.{idx}.conv_block'
        h = orc.conv(orc.pad_reflect(a, 1), sd[f'{p}.1.weight'], sd[f'{p}.1.bias'])
        h = relu(orc.instnorm(h))
        h = orc.conv(orc.pad_reflect(h, 1), sd[f'{p}.5.weight'], sd[f'{p}.5.bias'])
        a = tap(idx, a + orc.instnorm(h))
        idx += 1
    for _ in range(2):
        a = tap(idx, orc.blur_up(a))
        a = tap(idx + 1, orc.conv(a, sd[f'model.{idx + 1}.weight'], sd[f'model.{idx + 1}.bias'], pad=1))
        a = relu(orc.instnorm(a)); tap(idx + 2, a); tap(idx + 3, a)
        idx += 4
    a = tap(idx, orc.pad_reflect(a, 3))
    a = tap(idx + 1, orc.conv(a, sd[f'model.{idx + 1}.weight'], sd[f'model.{idx + 1}.bias']))
    fake = tap(idx + 2, np.tanh(a).astype(np.float32))
    return fake, feats


def unet(x, sd, n_enc, n_dec, prefix='unet_model.'):
    enc = [x]
    for i in range(n_enc):
        enc.append(leaky(orc.conv(enc[-1], sd[f'{prefix}downarm.{i}.main.weight'], sd[f'{prefix}downarm.{i}.main.bias'], stride=2, pad=1)))
    a = enc.pop()
    for i in range(n_enc):
        a = leaky(orc.conv(a, sd[f'{prefix}uparm.{i}.main.weight'], sd[f'{prefix}uparm.{i}.main.bias'], pad=1))
        a = np.concatenate([orc.upsample_nn(a), enc.pop()], axis=1)
    for i in range(n_dec - n_enc):
        a = leaky(orc.conv(a, sd[f'{prefix}extras.{i}.main.weight'], sd[f'{prefix}extras.{i}.main.bias'], pad=1))
    return a


def vxm_dense(source, target, sd, n_enc, n_dec, int_steps=7):
    """bidir VxmDense.forward: (y_source, y_target, pos_flow)."""
    x = unet(np.concatenate([source, target], axis=1), sd, n_enc, n_dec)
    flow = orc.conv(x, sd['flow.weight'], sd['flow.bias'], pad=1)
    pos = orc.resize_transform(flow, 2)
    neg = -pos
    pos, neg = orc.vecint(pos, int_steps), orc.vecint(neg, int_steps)
    pos, neg = orc.resize_transform(pos, 0.5), orc.resize_transform(neg, 0.5)
    return orc.warp(source, pos), orc.warp(target, neg), pos


def patch_sample(feat, ids, sd=None, mlp_id=0):
    """feat (B,C,H,W), ids (P,) -> (B*P, nc) L2-normalised rows."""
    B, C = feat.shape[:2]
    x = feat.transpose(0, 2, 3, 1).reshape(B, -1, C)[:, ids, :].reshape(-1, C).astype(np.float32)
    if sd is not None:
        x = relu(x @ sd[f'mlp_{mlp_id}.0.weight'].T + sd[f'mlp_{mlp_id}.0.bias'])
        x = (x @ sd[f'mlp_{mlp_id}.2.weight'].T + sd[f'mlp_{mlp_id}.2.bias']).astype(np.float32)
    norm = np.sqrt((x * x).sum(1, keepdims=True))
    return (x / (norm + np.float32(1e-7))).astype(np.float32)
