"""GPU parity: the sm_100a kernels (through the C ABI) against the C oracle and the reference's
golden vectors.  Integer / index outputs: bit-exact.  fp32 values: tolerances stated per test."""
import numpy as np
import pytest
import torch

import inputs as gi

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def L():
    import dfmir_b200.layers as layers
    return layers


@pytest.fixture(scope="module")
def LS():
    import dfmir_b200.losses as losses
    return losses


@pytest.mark.parametrize("case", gi.WARP_CASES, ids=[c[0] for c in gi.WARP_CASES])
def test_warp_indices_bit_exact(case, golden, orc, L):
    name, shape, sigma, seed = case
    g = golden("warp")
    flow = gi.flow(seed, 1, shape, sigma)
    ramps = gi.index_ramps(1, shape)
    # nearest mode: sampled integer index, against the reference's own output
    out, idx = L.warp_indices(cu(ramps), cu(flow), mode="nearest")
    assert np.array_equal(out.cpu().numpy().astype(np.int16), g[name + "/nearest_idx"])
    _, oidx = orc.warp(ramps, flow, mode="nearest", return_idx=True)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    # linear mode: floor() corner indices, bit-exact against the oracle
    img = gi.image(seed + 100, 1, shape)
    out, idx = L.warp_indices(cu(img), cu(flow), mode="bilinear")
    oout, oidx = orc.warp(img, flow, return_idx=True)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    np.testing.assert_allclose(out.cpu().numpy(), g[name + "/linear"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("case", gi.HALF_CASES, ids=[c[0] for c in gi.HALF_CASES])
def test_warp_half_integer_rounding(case, golden, L):
    name, shape, seed = case
    g = golden("warp")
    flow = gi.half_integer_flow(seed, 1, shape)
    out = L.warp(cu(gi.index_ramps(1, shape)), cu(flow), mode="nearest")
    assert np.array_equal(out.cpu().numpy().astype(np.int16), g[name + "/nearest_idx"])


def test_warp_matches_torch_cuda_grid_sample(L):
    """coord_mode=1 reproduces the CUDA ATen arithmetic (x * (1/(S-1))): compare nearest-mode
    indices with torch's own CUDA grid_sample driven by the reference's formula (layers.py:32-48)."""
    import torch.nn.functional as F
    for shape, sigma in (((256, 256), 1e-5), ((256, 256), 3.0), ((32, 40, 48), 1e-5), ((32, 40, 48), 3.0)):
        nd = len(shape)
        flow = cu(gi.flow(77, 1, shape, sigma))
        ramps = cu(gi.index_ramps(1, shape))
        grid = torch.stack(torch.meshgrid(*[torch.arange(s, device="cuda") for s in shape], indexing="ij")).float()[None]
        new_locs = grid + flow
        for i in range(nd):
            new_locs[:, i, ...] = 2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5)
        new_locs = new_locs.permute(0, *range(2, nd + 2), 1)[..., list(range(nd))[::-1]]
        ref = F.grid_sample(ramps, new_locs, align_corners=True, mode="nearest")
        got = L.warp(ramps, flow, mode="nearest", coord_mode=1)
        assert torch.equal(got, ref)


@pytest.mark.parametrize("name,shape,sigma,seed", [("b2d", (48, 64), 2.0, 41), ("b3d", (12, 16, 20), 1.5, 42)])
def test_warp_backward(name, shape, sigma, seed, golden, L):
    g = golden("warp_bwd")
    src = cu(gi.image(seed, 2, shape, C=2)).requires_grad_()
    flow = cu(gi.flow(seed + 1, 2, shape, sigma)).requires_grad_()
    gout = cu(gi.weights(seed + 2, (2, 2, *shape), 1.0))
    y = L.SpatialTransformer(shape).cuda()(src, flow)
    y.backward(gout)
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[name + "/out"], atol=2e-6)
    np.testing.assert_allclose(src.grad.cpu().numpy(), g[name + "/d_src"], atol=1e-5)
    np.testing.assert_allclose(flow.grad.cpu().numpy(), g[name + "/d_flow"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("name,shape,sigma,seed", [("v2d", (128, 128), 8.0, 51), ("v3d", (16, 20, 24), 4.0, 52),
                                                  ("v2d_tiny", (128, 128), 1e-3, 53)])
def test_vecint(name, shape, sigma, seed, golden, orc, L):
    g = golden("vecint_resize")
    vec_np = gi.smooth_field(gi.rng(seed), (2, len(shape), *shape), sigma)
    vec = cu(vec_np).requires_grad_()
    vi = L.VecInt(shape, 7).cuda()
    y = vi(vec)
    for got, want in ((y.detach().cpu().numpy(), g[name + "/out"]),
                      (y.detach().cpu().numpy(), orc.vecint(vec_np, 7))):
        err = np.abs(got - want)
        assert err.max() < 1e-3 and (err > 2e-5).mean() < 2e-3
    y.backward(cu(gi.weights(seed + 2, (2, len(shape), *shape), 1.0)))
    err = np.abs(vec.grad.cpu().numpy() - g[name + "/d_vec"])
    scale = np.abs(g[name + "/d_vec"]).max()
    assert err.max() < 2e-3 * scale and (err > 1e-4 * scale).mean() < 5e-3
    # both directions in one launch == two separate integrations
    pos, neg = vi.forward_bidir(vec.detach())
    assert torch.equal(pos, y.detach())
    err = np.abs(neg.cpu().numpy() - g[name + "/out_neg"])
    assert err.max() < 1e-3 and (err > 2e-5).mean() < 2e-3


@pytest.mark.parametrize("name,shape,seed", [("r2d", (64, 96), 61), ("r3d", (16, 20, 24), 62), ("r2d_odd", (37, 53), 63)])
def test_resize(name, shape, seed, golden, L):
    g = golden("vecint_resize")
    nd = len(shape)
    x = cu(gi.weights(seed, (2, nd, *shape), 1.0)).requires_grad_()
    down = L.ResizeTransform(2, nd)(x)
    np.testing.assert_allclose(down.detach().cpu().numpy(), g[name + "/down"], atol=1e-6)
    down.backward(cu(gi.weights(seed + 1, tuple(down.shape), 1.0)))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[name + "/down_dx"], atol=1e-5)
    x.grad = None
    up = L.ResizeTransform(0.5, nd)(x)
    np.testing.assert_allclose(up.detach().cpu().numpy(), g[name + "/up"], atol=2e-6)
    up.backward(cu(gi.weights(seed + 2, tuple(up.shape), 1.0)))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[name + "/up_dx"], atol=2e-5)


@pytest.mark.parametrize("name,shape,seed", [("n2d", (64, 64), 71), ("n3d", (24, 28, 32), 72), ("n2d_odd", (45, 70), 73)])
def test_ncc(name, shape, seed, golden, orc, LS):
    g = golden("losses")
    nd = len(shape)
    I = cu(gi.image_textured(seed, 2, shape)).requires_grad_()
    J = cu(gi.image_textured(seed + 1, 2, shape)).requires_grad_()
    crit = LS.NCC_Loss('cuda', kernel_var=[9] * nd, kernel_type='mean')
    loss = crit(I, J)
    assert abs(loss.item() - float(g[name + "/loss"])) <= 1e-4      # north_star: fp32 NCC within 1e-4
    assert abs(loss.item() - float(orc.ncc(I.detach().cpu().numpy(), J.detach().cpu().numpy())[0])) <= 1e-4
    loss.backward()
    for got, want in ((I.grad, g[name + "/dI"]), (J.grad, g[name + "/dJ"])):
        scale = np.abs(want).max()
        np.testing.assert_allclose(got.cpu().numpy(), want, atol=1e-2 * scale)
    I.grad = None
    mask = cu((gi.image(seed + 2, 2, shape) > -0.5).astype(np.float32))
    lm = crit(I, J, mask=mask)
    assert abs(lm.item() - float(g[name + "/loss_masked"])) <= 1e-4
    lm.backward()
    want = g[name + "/dI_masked"]
    np.testing.assert_allclose(I.grad.cpu().numpy(), want, atol=1e-2 * np.abs(want).max())
    assert abs(crit(I.detach(), I.detach()).item() - float(g[name + "/loss_self"])) <= 1e-4
    assert crit(I.detach(), J.detach(), mask=torch.zeros_like(mask)).item() == 0.0
    # vxm variant: -mean(cc)
    vl = LS.NCC([9] * nd).loss(I.detach(), J.detach())
    assert abs(vl.item() + float(g[name + "/cc"].mean())) <= 1e-4


@pytest.mark.parametrize("name,shape,seed", [("g2d", (64, 80), 81), ("g3d", (12, 16, 20), 82)])
def test_grad_loss(name, shape, seed, golden, LS):
    g = golden("losses")
    nd = len(shape)
    for pen, mult in (("l1", 2), ("l2", None)):
        x = cu(gi.weights(seed, (2, nd, *shape), 1.0)).requires_grad_()
        loss = LS.Grad_Loss(dim=nd, penalty=pen, loss_mult=mult)(x)
        want = float(g[f"{name}/{pen}"])
        assert abs(loss.item() - want) <= 2e-6 * abs(want)
        loss.backward()
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{name}/{pen}_dx"], atol=1e-8, rtol=1e-4)


def test_smoothing_and_l1(golden, LS):
    g = golden("losses")
    x = cu(gi.weights(91, (2, 2, 64, 80), 1.0)).requires_grad_()
    loss = LS.smooothing_loss(x)
    assert abs(loss.item() - float(g["smooth/loss"])) <= 2e-6 * abs(float(g["smooth/loss"]))
    loss.backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["smooth/dx"], atol=1e-9, rtol=1e-4)

    a = cu(gi.image(101, 2, (64, 64))).requires_grad_()
    b = cu(gi.image(102, 2, (64, 64))).requires_grad_()
    mask = (b > -0.95) + (a > -0.95)
    for loss in (LS.calculate_L1_loss(a, b, mask), LS.l1_threshold_masked(a, b, b, a, -0.95)):
        assert abs(loss.item() - float(g["l1/loss"])) <= 1e-6
        a.grad = None; b.grad = None
        loss.backward()
        np.testing.assert_allclose(a.grad.cpu().numpy(), g["l1/da"], atol=1e-9, rtol=1e-5)
        np.testing.assert_allclose(b.grad.cpu().numpy(), g["l1/db"], atol=1e-9, rtol=1e-5)
    assert LS.calculate_L1_loss(a.detach(), b.detach(), torch.zeros_like(mask)).item() == 0.0
    assert abs(LS.calculate_L1_loss(a.detach(), b.detach(), None).item() - float(g["l1/nomask"])) <= 1e-6


def test_full_size_properties(L, LS):
    """BASELINE sizes (3-D 128^3, B=2): size-independent properties instead of stored vectors."""
    shape = (128, 128, 128)
    img = cu(gi.image(7, 2, shape))
    st = L.SpatialTransformer(shape).cuda()
    # integer shifts are exact rolls in the interior
    flow = torch.zeros(2, 3, *shape, device="cuda")
    flow[:, 0] = 2.0; flow[:, 2] = -3.0
    out = st(img, flow)
    assert torch.allclose(out[:, :, :120, :, 8:], img[:, :, 2:122, :, 5:125], atol=2e-5)
    # linearity in src
    a, b = cu(gi.image(8, 2, shape)), cu(gi.image(9, 2, shape))
    f = cu(gi.flow(10, 2, shape, 3.0))
    assert torch.allclose(st(a + 2 * b, f), st(a, f) + 2 * st(b, f), atol=1e-5)
    # NCC(I, I) = -1, symmetric in its arguments, Grad of a constant field = 0
    ncc = LS.NCC_Loss('cuda', kernel_var=[9, 9, 9])
    ta, tb = cu(gi.image_textured(8, 2, shape, flat_bg=False)), cu(gi.image_textured(9, 2, shape, flat_bg=False))
    assert abs(ncc(ta, ta).item() + 1.0) < 1e-4          # every window textured: cc = 1 everywhere
    assert abs(ncc(ta, tb).item() - ncc(tb, ta).item()) < 1e-6
    assert LS.Grad_Loss(dim=3)(torch.full((2, 3, *shape), 0.25, device="cuda")).item() == 0.0
    # VecInt of a zero field is zero; of a constant field c is c (interior)
    vi = L.VecInt((64, 64, 64), 7).cuda()
    assert vi(torch.zeros(2, 3, 64, 64, 64, device="cuda")).abs().max().item() == 0.0
    c = torch.full((2, 3, 64, 64, 64), 0.5, device="cuda")
    assert torch.allclose(vi(c)[..., 8:-8, 8:-8, 8:-8], c[..., 8:-8, 8:-8, 8:-8], atol=1e-5)
