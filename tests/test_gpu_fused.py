"""The fused cooperative launch (integrate -> resize -> warp -> NCC + Grad, csrc/fused_reg.cu) against
(a) the chain of stand-alone kernels: flow and warped image bit-identical, losses to reduction order,
gradients to atomics order; (b) the C oracle; (c) at BASELINE size 128^3, through size-independent
properties (zero velocity = identity up to the reference's own fp32 round trip, NCC(I, I) = -1)."""
import numpy as np
import pytest
import torch

import inputs as gi

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def unfused(vel, moving, fixed, nd, nsteps, win):
    from dfmir_b200 import layers, losses
    half = list(vel.shape[2:])
    full = [2 * s for s in half]
    flow = layers.ResizeTransform(0.5, nd)(layers.VecInt(half, nsteps).cuda()(vel))
    warped = layers.SpatialTransformer(full).cuda()(moving, flow)
    ncc = losses.NCC_Loss('cuda', kernel_var=[win] * nd)(warped, fixed)
    grad = losses.Grad_Loss(dim=nd)(flow)
    return warped, flow, ncc, grad


@pytest.mark.parametrize("half,B,sigma", [((24, 32), 2, 1.5), ((12, 16, 20), 1, 1.0), ((16, 20, 24), 2, 4.0)])
def test_fused_matches_unfused_and_oracle(half, B, sigma, orc):
    from dfmir_b200 import integrate_warp_loss
    nd = len(half)
    full = tuple(2 * s for s in half)
    vel = gi.smooth_field(gi.rng(41 + nd), (B, nd, *half), sigma)
    moving = gi.image_textured(42, B, full, flat_bg=False)
    fixed = gi.image_textured(43, B, full, flat_bg=False)
    w1 = gi.weights(44, (B, 1, *full), 1.0)
    w2 = gi.weights(45, (B, nd, *full), 0.1)

    def run(fn):
        v = cu(vel).requires_grad_()
        m = cu(moving).requires_grad_()
        warped, flow, ncc, grad = fn(v, m, cu(fixed))
        loss = ncc + 0.5 * grad + (warped * cu(w1)).sum() * 1e-3 + (flow * cu(w2)).sum() * 1e-3
        loss.backward()
        return warped.detach(), flow.detach(), float(ncc), float(grad), v.grad, m.grad

    fw, ff, fn_, fg, fdv, fdm = run(lambda v, m, f: integrate_warp_loss(v, m, f, nsteps=7, win=9))
    uw, uf, un, ug, udv, udm = run(lambda v, m, f: unfused(v, m, f, nd, 7, 9))
    assert torch.equal(ff, uf), "integrated full-resolution flow differs from the stand-alone kernels"
    assert torch.equal(fw, uw), "warped image differs from the stand-alone kernels"
    assert abs(fn_ - un) <= 1e-6 and abs(fg - ug) <= 1e-6 * max(1.0, abs(ug)), (fn_, un, fg, ug)
    for a, b, what in ((fdv, udv, "d vel"), (fdm, udm, "d moving")):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-5 * max(scale, 1e-12), what
    # against the C oracle (same checks as smoke())
    o_flow = orc.resize_transform(orc.vecint(vel, 7), 0.5)
    o_warped = orc.warp(moving, o_flow)
    # 7 squarings + a resize of fp32 interpolation (fused multiply-adds here, separate roundings in the C oracle)
    np.testing.assert_allclose(ff.cpu().numpy(), o_flow, atol=1e-5)
    np.testing.assert_allclose(fw.cpu().numpy(), o_warped, atol=2e-5)
    assert abs(fn_ - float(orc.ncc(o_warped, fixed)[0])) <= 1e-4
    o_grad = orc.grad_loss(o_flow, 2)
    assert abs(fg - o_grad) <= 1e-5 * max(1.0, abs(o_grad))


def test_fused_full_size_properties():
    """BASELINE configs[2] size (128^3, batch 2): properties that need no oracle."""
    from dfmir_b200 import integrate_warp_loss
    B, half = 2, (64, 64, 64)
    full = (128, 128, 128)
    img = cu(gi.image_textured(51, B, full, flat_bg=False))
    zero = torch.zeros((B, 3, *half), device="cuda")
    warped, flow, ncc, grad = integrate_warp_loss(zero, img, img, nsteps=7, win=9)
    assert float(flow.abs().max()) == 0.0 and float(grad) == 0.0
    # zero flow is the identity only up to the reference's normalise/unnormalise round trip (1.5e-6 per
    # axis in 2-D at unit intensity; three axes here)
    assert float((warped - img).abs().max()) <= 2e-5
    assert abs(float(ncc) + 1.0) <= 1e-4
    # a constant integer shift along x: the integrated flow of a constant velocity is that constant
    vel = torch.zeros((B, 3, *half), device="cuda")
    vel[:, 2] = 1.0          # half-resolution units; the full-resolution flow is 2 voxels
    warped, flow, ncc, grad = integrate_warp_loss(vel, img, img, nsteps=7, win=9)
    inner = flow[:, 2, 8:-8, 8:-8, 8:-8]
    assert float((inner - 2.0).abs().max()) <= 1e-4 and float(flow[:, :2].abs().max()) <= 1e-6
    shifted = torch.roll(img, shifts=-2, dims=4)
    assert float((warped[..., 8:-8, 8:-8, 8:-8] - shifted[..., 8:-8, 8:-8, 8:-8]).abs().max()) <= 5e-4
    assert float(ncc) > -1.0 + 1e-3      # a shifted image correlates less than the image itself
