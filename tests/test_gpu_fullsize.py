"""BASELINE configs[3] size (3-D 160 x 192 x 160, one pair per GPU): checks that do not need the CPU oracle to
finish at this size.

* warp indices against ATen's own CUDA grid_sample driven by the reference's coordinate formula
  (models/voxelmorph/torchvoxelmorph/layers.py:30-48): nearest-mode samples of index ramps must be equal
  (bit-identical deformation-field indices), linear mode to interpolation rounding;
* NCC / Grad against the reference formulas (util/losses.py:176-261, 81-130) evaluated in float64 with torch
  pooling ops: |NCC - ref| <= 1e-4;
* the fused launch against the chain of stand-alone kernels (flow and warped volume bit-identical);
* one VoxelMorph-3D training step with the reference's default features (vxm/networks.py:9-14): finite losses and
  gradients, fused == unfused forward, tensor-core gradients aligned with the fp32 engine's.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import inputs as gi

pytestmark = pytest.mark.gpu

SHAPE = (160, 192, 160)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def aten_warp(src, flow, mode):
    """SpatialTransformer.forward of the reference, op for op, on ATen CUDA."""
    shape = flow.shape[2:]
    nd = len(shape)
    grid = torch.stack(torch.meshgrid(*[torch.arange(s, device=flow.device) for s in shape], indexing="ij")).float()[None]
    new_locs = grid + flow
    for i in range(nd):
        new_locs[:, i, ...] = 2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5)
    new_locs = new_locs.permute(0, *range(2, nd + 2), 1)[..., list(range(nd))[::-1]]
    return F.grid_sample(src, new_locs, align_corners=True, mode=mode)


@pytest.mark.parametrize("sigma", [1e-5, 3.0, 20.0])
def test_warp_indices_full_size(sigma):
    import dfmir_b200.layers as L
    flow = cu(gi.flow(301, 1, SHAPE, sigma))
    ramps = cu(gi.index_ramps(1, SHAPE))
    got = L.warp(ramps, flow, mode="nearest", coord_mode=1)
    assert torch.equal(got, aten_warp(ramps, flow, "nearest"))
    img = cu(gi.image_textured(302, 1, SHAPE, flat_bg=False))
    got = L.warp(img, flow, mode="bilinear", coord_mode=1)
    ref = aten_warp(img, flow, "bilinear")
    assert float((got - ref).abs().max()) <= 4e-6


def box_sum64(x, win):
    nd = x.dim() - 2
    pool = getattr(F, "avg_pool%dd" % nd)
    return pool(x, win, stride=1, padding=win // 2, count_include_pad=True) * float(win ** nd)


def ncc_ref64(I, J, win=9, eps=1e-5):
    """cc of util/losses.py:176-238 in float64; returns -sqrt(mean(cc))."""
    I, J = I.double(), J.double()
    W = float(win ** (I.dim() - 2))
    Is, Js, I2, J2, IJ = (box_sum64(t, win) for t in (I, J, I * I, J * J, I * J))
    uI, uJ = Is / W, Js / W
    cross = IJ - uJ * Is - uI * Js + uI * uJ * W
    Iv = I2 - 2 * uI * Is + uI * uI * W
    Jv = J2 - 2 * uJ * Js + uJ * uJ * W
    cc = cross * cross / (Iv * Jv + eps)
    return -torch.sqrt(cc.mean())


def grad_ref64(flow):
    """Grad_Loss(dim=3, penalty='l2') of util/losses.py:81-130 in float64."""
    f = flow.double()
    d = [(f[:, :, 1:] - f[:, :, :-1]), (f[:, :, :, 1:] - f[:, :, :, :-1]), (f[..., 1:] - f[..., :-1])]
    return sum((t * t).mean() for t in d) / 3.0


def test_ncc_grad_full_size():
    from dfmir_b200 import losses
    a = cu(gi.image_textured(311, 1, SHAPE, flat_bg=False))
    b = cu(gi.image_textured(312, 1, SHAPE, flat_bg=False))
    got = float(losses.NCC_Loss('cuda', kernel_var=[9, 9, 9])(a, b))
    assert abs(got - float(ncc_ref64(a, b))) <= 1e-4
    assert abs(float(losses.NCC_Loss('cuda', kernel_var=[9, 9, 9])(a, a)) + 1.0) <= 1e-4
    flow = cu(gi.flow(313, 1, SHAPE, 3.0))
    g = float(losses.Grad_Loss(dim=3)(flow))
    r = float(grad_ref64(flow))
    assert abs(g - r) <= 1e-5 * max(1.0, abs(r))


def test_fused_full_size_bit_identical():
    from dfmir_b200 import integrate_warp_loss, layers, losses
    half = tuple(s // 2 for s in SHAPE)
    vel = cu(gi.smooth_field(gi.rng(321), (1, 3, *half), 2.0))
    moving = cu(gi.image_textured(322, 1, SHAPE, flat_bg=False))
    fixed = cu(gi.image_textured(323, 1, SHAPE, flat_bg=False))
    warped, flow, ncc, grad = integrate_warp_loss(vel, moving, fixed, nsteps=7, win=9)
    uflow = layers.ResizeTransform(0.5, 3)(layers.VecInt(list(half), 7).cuda()(vel))
    uwarped = layers.SpatialTransformer(list(SHAPE)).cuda()(moving, uflow)
    assert torch.equal(flow, uflow) and torch.equal(warped, uwarped)
    assert abs(float(ncc) - float(ncc_ref64(uwarped, fixed))) <= 1e-4
    r = float(grad_ref64(uflow))
    assert abs(float(grad) - r) <= 1e-5 * max(1.0, abs(r))
    assert abs(float(ncc) - float(losses.NCC_Loss('cuda', kernel_var=[9, 9, 9])(uwarped, fixed))) <= 1e-6


def test_vxm_default_features_step_full_size(monkeypatch):
    """One registration step at 160 x 192 x 160 with the reference's default U-Net features."""
    import dfmir_b200.functional as Fn
    from dfmir_b200 import losses, vxm
    torch.manual_seed(5)
    R = vxm.VxmDense(SHAPE, int_steps=7, bidir=False).cuda()
    with torch.no_grad():
        R.flow.weight.mul_(1e4)          # N(0, 1e-5) initial flow head: scale it so the warp is not the identity
    src = cu(gi.image_textured(331, 1, SHAPE, flat_bg=False))
    tgt = cu(gi.image_textured(332, 1, SHAPE, flat_bg=False))

    def step(engine, fused):
        monkeypatch.setattr(Fn, "CONV_ENGINE", engine)
        for p in R.parameters():
            p.grad = None
        if fused:
            y, flow, ncc, grad = R.forward_with_losses(src, tgt, win=9)
        else:
            y, flow = R(src, tgt, registration=True)
            ncc = losses.NCC_Loss('cuda', kernel_var=[9, 9, 9])(y, tgt)
            grad = losses.Grad_Loss(dim=3)(flow)
        (ncc + 0.02 * grad).backward()
        g = torch.cat([p.grad.flatten() for p in R.parameters() if p.grad is not None])
        return y.detach(), flow.detach(), float(ncc.detach()), float(grad.detach()), g

    default = Fn.CONV_ENGINE
    y1, f1, n1, g1, gr1 = step(default, True)
    y2, f2, n2, g2, gr2 = step(default, False)
    assert torch.isfinite(gr1).all() and np.isfinite([n1, g1]).all()
    assert float(f1.abs().max()) > 1e-3, "the test needs a non-trivial deformation"
    assert torch.equal(f1, f2) and torch.equal(y1, y2)
    assert abs(n1 - n2) <= 1e-6 and abs(g1 - g2) <= 1e-6 * max(1.0, abs(g2))
    cos = float(F.cosine_similarity(gr1, gr2, dim=0))
    assert cos >= 0.9999, cos
    if default != "simt":
        # TF32-operand tensor-core engine against the fp32 CUDA-core engine: same step, gradients aligned
        y3, f3, n3, g3, gr3 = step("simt", True)
        assert abs(n1 - n3) <= 1e-3 and float((f1 - f3).abs().max()) <= 5e-2 * float(f3.abs().max())
        cos = float(F.cosine_similarity(gr1, gr3, dim=0))
        assert cos >= 0.98, cos
