"""CPU port of the reference's training step on the library the reference itself runs on
(PyTorch fp32 CPU ops: F.conv2d / F.instance_norm / F.grid_sample / F.interpolate / torch.bmm),
restated functionally over plain state-dicts.  TEST INFRASTRUCTURE ONLY: used by tests/ as the
autograd checker at sizes the C oracle cannot reach, and by bench.py's cpu_baseline /
`--impl reference` legs as "the reference's CPU PyTorch path" on the GPU box (where
/root/reference does not exist).  Pinned against the reference's own step in
tests/test_oracle_nets.py::test_torch_port_step (golden: tests/golden/step.npz).

Follows (paths relative to the reference repo):
    resnet_generator   models/networks.py:956-1051, ResnetBlock :1164-1221, Downsample :37-60, Upsample :73-93
    patch_sample       models/networks.py:597-624, Normalize :493-502
    patchnce           models/patchnce.py:14-55
    spatial_transform  models/voxelmorph/torchvoxelmorph/layers.py:30-48
    vec_int / resize   models/voxelmorph/torchvoxelmorph/layers.py:64-68, 85-97
    vxm_dense          models/voxelmorph/torchvoxelmorph/networks.py:88-106, 1102-1145
    step               models/registration_model.py:138-171, 213-263
    ncc_loss / grad_loss  util/losses.py:183-261 (box filters as dense convolutions, like the reference), :81-130
    Step3D             VxmDense-3D + NCC_Loss + Grad_Loss forward / backward + Adam: the 3-D workloads of BASELINE
                       configs[2..4] (BASELINE.md section 2 times exactly this composition of the reference)
"""
import numpy as np
import torch
import torch.nn.functional as F

_BLUR3 = torch.tensor([1., 2., 1.])
_BLUR4 = torch.tensor([1., 3., 3., 1.])


# ---- TF32 operand emulation (for checking the tcgen05 engine) -------------------------------------
# The tensor core reads fp32 operands and keeps sign, exponent and the top 10 mantissa bits
# (truncation: the low 13 bits are ignored); products are exact, accumulation is fp32.  With
# TF32_EMULATION set, the convolutions the tensor-core engine covers (generator: Cin and Cout in {64,128,256};
# registration U-Net: Cin and Cout >= 16, 2-D / 3-D, stride 1 / 2) run forward / data-gradient / weight-gradient
# on truncated operands, in the given accumulation dtype.  None (default) = plain F.conv{2,3}d, the reference's
# CPU arithmetic.
TF32_EMULATION = None      # None | "trunc" | "rna"


def tf32_round(t, mode="trunc"):
    f = t.detach().to(torch.float32).contiguous()
    i = f.view(torch.int32)
    if mode == "rna":       # round to nearest, ties away from zero (cvt.rna.tf32.f32)
        i = i + 0x1000
    i = i & ~0x1FFF
    return i.view(torch.float32).to(t.dtype)


class _Tf32Conv(torch.autograd.Function):
    """conv{2,3}d on TF32-truncated operands, all three products (forward, data gradient, weight gradient)."""

    @staticmethod
    def forward(ctx, x, w, b, padding, stride):
        ctx.save_for_backward(x, w)
        ctx.padding, ctx.stride = padding, stride
        conv = F.conv2d if x.dim() == 4 else F.conv3d
        return conv(tf32_round(x, TF32_EMULATION), tf32_round(w, TF32_EMULATION), b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gq, xq, wq = tf32_round(gy, TF32_EMULATION), tf32_round(x, TF32_EMULATION), tf32_round(w, TF32_EMULATION)
        g = torch.nn.grad
        d_in, d_w = (g.conv2d_input, g.conv2d_weight) if x.dim() == 4 else (g.conv3d_input, g.conv3d_weight)
        gx = d_in(x.shape, wq, gq, stride=ctx.stride, padding=ctx.padding)
        gw = d_w(xq, w.shape, gq, stride=ctx.stride, padding=ctx.padding)
        return gx, gw, gy.sum(dim=[0] + list(range(2, gy.dim()))), None, None


def conv2d(x, w, b, padding=0):
    if TF32_EMULATION and w.shape[0] in (64, 128, 256) and w.shape[1] in (64, 128, 256):
        return _Tf32Conv.apply(x, w, b, padding, 1)
    return F.conv2d(x, w, b, padding=padding)


def unet_conv(x, w, b, stride=1, padding=1):
    """A U-Net convolution (vxm/networks.py:1506-1521).  Under TF32_EMULATION the layers the tcgen05 engine covers
    (reduction-side channels >= 16: every layer but the 2-channel input layer and the nd-channel flow head, stride 2
    included - the engine runs those in space-to-depth form on the same operands) use truncated operands."""
    if TF32_EMULATION and w.shape[0] >= 16 and w.shape[1] >= 16:
        return _Tf32Conv.apply(x, w, b, padding, stride)
    conv = F.conv2d if x.dim() == 4 else F.conv3d
    return conv(x, w, b, stride=stride, padding=padding)


def _blur_filt(a, C, scale=1.0):
    f = a[:, None] * a[None, :]
    return (f / f.sum() * scale)[None, None].repeat(C, 1, 1, 1)


def blur_down(x):
    C = x.shape[1]
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), _blur_filt(_BLUR3, C).to(x), stride=2, groups=C)


def blur_up(x):
    C = x.shape[1]
    y = F.conv_transpose2d(F.pad(x, (1, 1, 1, 1), mode='replicate'), _blur_filt(_BLUR4, C, 4.0).to(x), stride=2,
                           padding=2, groups=C)
    return y[:, :, 1:, 1:][:, :, :-1, :-1]


def resnet_generator(x, sd, n_blocks, layers=(), encode_only=False, p='model.'):
    feats = []
    last = layers[-1] if len(layers) else None

    class Stop(Exception):
        pass

    def tap(i, v):
        if i in layers:
            feats.append(v)
        if encode_only and i == last:
            raise Stop()
        return v

    def inr(v):
        return F.relu(F.instance_norm(v))

    try:
        a = tap(0, F.pad(x, (3,) * 4, mode='reflect'))
        a = tap(1, conv2d(a, sd[p + '1.weight'], sd[p + '1.bias']))
        a = inr(a); tap(2, a); tap(3, a)
        idx = 4
        for _ in range(2):
            a = tap(idx, conv2d(a, sd[f'{p}{idx}.weight'], sd[f'{p}{idx}.bias'], padding=1))
            a = inr(a); tap(idx + 1, a); tap(idx + 2, a)
            a = tap(idx + 3, blur_down(a))
            idx += 4
        for _ in range(n_blocks):
            q = f'{p}{idx}.conv_block.'
            h = conv2d(F.pad(a, (1,) * 4, mode='reflect'), sd[q + '1.weight'], sd[q + '1.bias'])
            h = inr(h)
            h = conv2d(F.pad(h, (1,) * 4, mode='reflect'), sd[q + '5.weight'], sd[q + '5.bias'])
            a = tap(idx, a + F.instance_norm(h))
            idx += 1
        for _ in range(2):
            a = tap(idx, blur_up(a))
            a = tap(idx + 1, conv2d(a, sd[f'{p}{idx + 1}.weight'], sd[f'{p}{idx + 1}.bias'], padding=1))
            a = inr(a); tap(idx + 2, a); tap(idx + 3, a)
            idx += 4
        a = tap(idx, F.pad(a, (3,) * 4, mode='reflect'))
        a = tap(idx + 1, conv2d(a, sd[f'{p}{idx + 1}.weight'], sd[f'{p}{idx + 1}.bias']))
        a = tap(idx + 2, torch.tanh(a))
    except Stop:
        return feats
    return (a, feats) if len(layers) else a


def patch_sample(feats, sd, num_patches, patch_ids=None):
    out, ids = [], []
    for i, feat in enumerate(feats):
        B, C, H, W = feat.shape
        fr = feat.permute(0, 2, 3, 1).flatten(1, 2)
        pid = patch_ids[i] if patch_ids is not None else torch.randperm(H * W, device=feat.device)[:min(num_patches, H * W)]
        x = fr[:, pid, :].flatten(0, 1)
        if sd is not None:
            x = F.linear(F.relu(F.linear(x, sd[f'mlp_{i}.0.weight'], sd[f'mlp_{i}.0.bias'])), sd[f'mlp_{i}.2.weight'],
                         sd[f'mlp_{i}.2.bias'])
        out.append(x / (x.pow(2).sum(1, keepdim=True).pow(0.5) + 1e-7))
        ids.append(pid)
    return out, ids


def patchnce(q, k, batch, T=0.07):
    k = k.detach()
    n, dim = q.shape
    l_pos = torch.bmm(q.view(n, 1, -1), k.view(n, -1, 1)).view(n, 1)
    qb, kb = q.view(batch, -1, dim), k.view(batch, -1, dim)
    P = qb.size(1)
    l_neg = torch.bmm(qb, kb.transpose(2, 1))
    l_neg = l_neg.masked_fill(torch.eye(P, dtype=torch.bool, device=q.device)[None], -10.0).view(-1, P)
    out = torch.cat((l_pos, l_neg), dim=1) / T
    return F.cross_entropy(out, torch.zeros(n, dtype=torch.long, device=q.device), reduction='none')


def spatial_transform(src, flow, mode='bilinear'):
    shape = flow.shape[2:]
    grid = torch.stack(torch.meshgrid(*[torch.arange(0, s) for s in shape], indexing='ij')).unsqueeze(0).float().to(flow)
    new_locs = grid + flow
    for i in range(len(shape)):
        new_locs[:, i, ...] = 2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5)
    if len(shape) == 2:
        new_locs = new_locs.permute(0, 2, 3, 1)[..., [1, 0]]
    else:
        new_locs = new_locs.permute(0, 2, 3, 4, 1)[..., [2, 1, 0]]
    return F.grid_sample(src, new_locs, align_corners=True, mode=mode)


def vec_int(vec, nsteps):
    vec = vec * (1.0 / (2 ** nsteps))
    for _ in range(nsteps):
        vec = vec + spatial_transform(vec, vec)
    return vec


def resize_transform(x, vel_resize):
    factor = 1.0 / vel_resize
    mode = 'bilinear' if x.dim() == 4 else 'trilinear'
    if factor < 1:
        return factor * F.interpolate(x, align_corners=True, scale_factor=factor, mode=mode)
    if factor > 1:
        return F.interpolate(factor * x, align_corners=True, scale_factor=factor, mode=mode)
    return x


def unet(x, sd, n_enc, n_dec, p='unet_model.'):
    conv = unet_conv
    enc = [x]
    for i in range(n_enc):
        enc.append(F.leaky_relu(conv(enc[-1], sd[f'{p}downarm.{i}.main.weight'], sd[f'{p}downarm.{i}.main.bias'], stride=2, padding=1), 0.2))
    a = enc.pop()
    for i in range(n_enc):
        a = F.leaky_relu(conv(a, sd[f'{p}uparm.{i}.main.weight'], sd[f'{p}uparm.{i}.main.bias'], padding=1), 0.2)
        a = torch.cat([F.interpolate(a, scale_factor=2, mode='nearest'), enc.pop()], dim=1)
    for i in range(n_dec - n_enc):
        a = F.leaky_relu(conv(a, sd[f'{p}extras.{i}.main.weight'], sd[f'{p}extras.{i}.main.bias'], padding=1), 0.2)
    return a


def vxm_dense(source, target, sd, n_enc, n_dec, int_steps=7):
    conv = F.conv2d if source.dim() == 4 else F.conv3d
    x = unet(torch.cat([source, target], dim=1), sd, n_enc, n_dec)
    pos = resize_transform(conv(x, sd['flow.weight'], sd['flow.bias'], padding=1), 2)
    neg = -pos
    pos, neg = vec_int(pos, int_steps), vec_int(neg, int_steps)
    pos, neg = resize_transform(pos, 0.5), resize_transform(neg, 0.5)
    return spatial_transform(source, pos), spatial_transform(target, neg), pos


def masked_l1(src, tgt, mask):
    diff = torch.abs(src - tgt)
    if torch.sum(mask) == 0:
        return torch.tensor(0)
    return (1 / torch.sum(mask)) * torch.sum(diff * mask)


def smoothing(y):
    dy = torch.abs(y[:, :, 1:, :] - y[:, :, :-1, :])
    dx = torch.abs(y[:, :, :, 1:] - y[:, :, :, :-1])
    return (torch.mean(dx * dx) + torch.mean(dy * dy)) / 2.0


def ncc_loss(prediction, target, win=9, eps=1e-5, mask=None):
    """NCC_Loss.forward with the 'mean' kernel: five box sums as dense conv{2,3}d with a ones filter, zero padding
    win // 2; cc = cross^2 / (I_var * J_var + eps); -sqrt(mean(cc))  (util/losses.py:183-261)."""
    nd = prediction.dim() - 2
    conv = getattr(F, 'conv%dd' % nd)
    filt = torch.ones([1, 1] + [win] * nd, dtype=prediction.dtype, device=prediction.device)
    pad = win // 2
    I, J = prediction, target
    I_sum, J_sum = conv(I, filt, padding=pad), conv(J, filt, padding=pad)
    I2_sum, J2_sum, IJ_sum = conv(I * I, filt, padding=pad), conv(J * J, filt, padding=pad), conv(I * J, filt, padding=pad)
    win_size = torch.sum(filt)
    u_I, u_J = I_sum / win_size, J_sum / win_size
    cross = IJ_sum - u_J * I_sum - u_I * J_sum + u_I * u_J * win_size
    I_var = I2_sum - 2 * u_I * I_sum + u_I * u_I * win_size
    J_var = J2_sum - 2 * u_J * J_sum + u_J * u_J * win_size
    cc = cross * cross / (I_var * J_var + eps)
    if mask is None:
        return -1.0 * torch.sqrt(torch.mean(cc))
    if torch.sum(mask) == 0:
        return torch.tensor(0)
    return -1.0 * torch.sqrt((1 / torch.sum(mask)) * torch.sum(cc * mask))


def grad_loss(prediction, penalty='l2'):
    """Grad_Loss(dim = nd): mean |forward difference|^p along each spatial axis, averaged over the axes
    (util/losses.py:81-130)."""
    nd = prediction.dim() - 2
    total = 0.0
    for ax in range(2, 2 + nd):
        n = prediction.shape[ax]
        d = torch.abs(prediction.narrow(ax, 1, n - 1) - prediction.narrow(ax, 0, n - 1))
        total = total + torch.mean(d * d if penalty == 'l2' else d)
    return total / float(nd)


def random_state_dict_r3d(feats, seed=0):
    """Random-initialised VxmDense-3D weights (PyTorch conv defaults + N(0, 1e-5) flow head) for timing runs."""
    g = torch.Generator().manual_seed(seed)
    enc, dec = feats

    def kaiming_u(shape):
        bound = (1.0 / (shape[1] * int(np.prod(shape[2:])))) ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    R, prev = {}, 2
    for i, nf in enumerate(enc):
        R[f'unet_model.downarm.{i}.main.weight'] = kaiming_u((nf, prev, 3, 3, 3)); R[f'unet_model.downarm.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    hist = list(reversed(enc))
    for i, nf in enumerate(dec[:len(enc)]):
        ch = prev + hist[i] if i > 0 else prev
        R[f'unet_model.uparm.{i}.main.weight'] = kaiming_u((nf, ch, 3, 3, 3)); R[f'unet_model.uparm.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    prev += 2
    for i, nf in enumerate(dec[len(enc):]):
        R[f'unet_model.extras.{i}.main.weight'] = kaiming_u((nf, prev, 3, 3, 3)); R[f'unet_model.extras.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    R['flow.weight'] = torch.randn((3, prev, 3, 3, 3), generator=g) * 1e-5
    R['flow.bias'] = torch.zeros(3)
    return R


class Step3D:
    """One 3-D registration training step the way the reference composes it:
    VxmDense(bidir=False).forward -> NCC_Loss([win]^3)(y_source, target) + lam * Grad_Loss(dim=3)(flow) -> backward -> Adam."""

    def __init__(self, sdR, feats, int_steps=7, win=9, lam=0.02, lr=2e-4, betas=(0.5, 0.999)):
        self.P = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith('.grid')) for k, v in sdR.items()}
        self.levels = (len(feats[0]), len(feats[1]))
        self.int_steps, self.win, self.lam = int_steps, win, lam
        self.opt = torch.optim.Adam([v for v in self.P.values() if v.requires_grad], lr=lr, betas=betas)
        self.losses = {}

    def forward(self, source, target):
        x = unet(torch.cat([source, target], dim=1), self.P, *self.levels)
        pos = resize_transform(F.conv3d(x, self.P['flow.weight'], self.P['flow.bias'], padding=1), 2)
        pos = resize_transform(vec_int(pos, self.int_steps), 0.5)
        return spatial_transform(source, pos), pos

    def step(self, source, target):
        self.opt.zero_grad()
        y, flow = self.forward(source, target)
        ncc = ncc_loss(y, target, self.win)
        grad = grad_loss(flow)
        (ncc + self.lam * grad).backward()
        self.opt.step()
        self.losses = {'ncc': float(ncc.detach()), 'grad': float(grad.detach())}
        self.y, self.flow = y.detach(), flow.detach()
        return self.losses


class Step:
    """REGISTRATIONModel.optimize_parameters over three state-dicts (G, F, R) with three Adams."""

    def __init__(self, sdG, sdF, sdR, n_blocks=9, batch_size=1, nce_layers=(0, 4, 8, 12, 16), num_patches=256,
                 nce_T=0.07, lambda_NCE=0.25, lr=2e-4, betas=(0.5, 0.999), r_levels=(6, 7), dvf_image=None):
        self.P = {n: {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(('.filt', '.grid')))
                      for k, v in sd.items()} for n, sd in (('G', sdG), ('F', sdF), ('R', sdR))}
        self.n_blocks, self.B, self.layers = n_blocks, batch_size, list(nce_layers)
        self.num_patches, self.T, self.lam, self.r_levels = num_patches, nce_T, lambda_NCE, r_levels
        self.opt = {n: torch.optim.Adam([v for v in d.values() if v.requires_grad], lr=lr, betas=betas)
                    for n, d in self.P.items()}
        self.dvf_image = dvf_image
        self.losses = {}

    def nce(self, src, tgt):
        G, Fs = self.P['G'], self.P['F']
        fq = resnet_generator(tgt, G, self.n_blocks, self.layers, encode_only=True)
        fk = resnet_generator(src, G, self.n_blocks, self.layers, encode_only=True)
        k_pool, ids = patch_sample(fk, Fs, self.num_patches, None)
        q_pool, _ = patch_sample(fq, Fs, self.num_patches, ids)
        total = 0.0
        for q, k in zip(q_pool, k_pool):
            total = total + (patchnce(q, k, self.B, self.T) * self.lam).mean()
        return total / len(self.layers)

    def step(self, real_A, real_B):
        B = real_A.shape[0]
        fake = resnet_generator(torch.cat((real_A, real_B), 0), self.P['G'], self.n_blocks)
        fake_B, idt_B = fake[:B], fake[B:]
        regA, _, pos_flow = vxm_dense(real_A, real_B, self.P['R'], *self.r_levels)
        registered = spatial_transform(fake_B, pos_flow)
        dvf = spatial_transform(self.dvf_image, pos_flow) if self.dvf_image is not None else None
        for o in self.opt.values():
            o.zero_grad()
        loss_NCE = self.nce(real_A, fake_B)
        loss_NCE_Y = self.nce(real_B, idt_B)
        loss_G = (loss_NCE + loss_NCE_Y) * 0.5
        mask = (real_B > -0.95) + (registered > -0.95)
        mask2 = (idt_B > -0.95) + (registered > -0.95)
        loss_local = self.nce(real_B, regA) * 0.25
        loss_R = masked_l1(registered, real_B, mask) + masked_l1(idt_B, registered, mask2) + loss_local
        loss_smooth = smoothing(pos_flow) * 0.20
        (loss_R + loss_G + loss_smooth).backward()
        for o in self.opt.values():
            o.step()
        self.losses = {'G': float(loss_G), 'NCE': float(loss_NCE), 'R': float(loss_R), 'smooth': float(loss_smooth),
                       'local': float(loss_local), 'NCE_Y': float(loss_NCE_Y)}
        self.visuals = {'fake_B': fake_B, 'idt_B': idt_B, 'registered': registered, 'regA': regA, 'dvf': dvf}
        return self.losses


def random_state_dicts(ngf=64, n_blocks=9, netF_nc=256, crop=256, seed=0,
                       r_feats=((16, 32, 32, 64, 64, 64), (64, 64, 64, 32, 32, 32, 16))):
    """Random-initialised weights of the reference architecture (xavier 0.02 for G/F, PyTorch conv
    defaults + N(0,1e-5) flow head for R): for timing runs, where no checkpoint exists."""
    g = torch.Generator().manual_seed(seed)

    def xavier(shape, gain=0.02):
        rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
        std = gain * (2.0 / (shape[1] * rf + shape[0] * rf)) ** 0.5
        return torch.randn(shape, generator=g) * std

    def kaiming_u(shape):
        fan_in = shape[1] * int(np.prod(shape[2:]))
        bound = (1.0 / fan_in) ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    G = {}

    def addc(name, co, ci, k):
        G[name + '.weight'] = xavier((co, ci, k, k)); G[name + '.bias'] = torch.zeros(co)
    addc('model.1', ngf, 1, 7); addc('model.4', ngf * 2, ngf, 3); addc('model.8', ngf * 4, ngf * 2, 3)
    idx = 12
    for _ in range(n_blocks):
        addc(f'model.{idx}.conv_block.1', ngf * 4, ngf * 4, 3); addc(f'model.{idx}.conv_block.5', ngf * 4, ngf * 4, 3)
        idx += 1
    addc(f'model.{idx + 1}', ngf * 2, ngf * 4, 3); addc(f'model.{idx + 5}', ngf, ngf * 2, 3); addc(f'model.{idx + 9}', 1, ngf, 7)
    Fd = {}
    for i, c in enumerate([1, ngf * 2, ngf * 4, ngf * 4, ngf * 4]):
        Fd[f'mlp_{i}.0.weight'] = xavier((netF_nc, c)); Fd[f'mlp_{i}.0.bias'] = torch.zeros(netF_nc)
        Fd[f'mlp_{i}.2.weight'] = xavier((netF_nc, netF_nc)); Fd[f'mlp_{i}.2.bias'] = torch.zeros(netF_nc)
    R = {}
    enc, dec = r_feats
    prev = 2
    for i, nf in enumerate(enc):
        R[f'unet_model.downarm.{i}.main.weight'] = kaiming_u((nf, prev, 3, 3)); R[f'unet_model.downarm.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    hist = list(reversed(enc))
    for i, nf in enumerate(dec[:len(enc)]):
        ch = prev + hist[i] if i > 0 else prev
        R[f'unet_model.uparm.{i}.main.weight'] = kaiming_u((nf, ch, 3, 3)); R[f'unet_model.uparm.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    prev += 2
    for i, nf in enumerate(dec[len(enc):]):
        R[f'unet_model.extras.{i}.main.weight'] = kaiming_u((nf, prev, 3, 3)); R[f'unet_model.extras.{i}.main.bias'] = torch.zeros(nf)
        prev = nf
    R['flow.weight'] = torch.randn((2, prev, 3, 3), generator=g) * 1e-5
    R['flow.bias'] = torch.zeros(2)
    return G, Fd, R
