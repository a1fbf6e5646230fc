"""CPU: the C oracle (oracle/dfmir_oracle.c) against golden vectors produced by the reference
itself (oracle/gen_golden.py).  Indices and normalised coordinates: bit-exact.  Values: fp32
tolerances stated per test."""
import numpy as np
import pytest

import inputs as gi


@pytest.mark.parametrize("case", gi.WARP_CASES, ids=[c[0] for c in gi.WARP_CASES])
def test_warp_indices_bit_exact(case, golden, orc):
    name, shape, sigma, seed = case
    g = golden("warp")
    flow = gi.flow(seed, 1, shape, sigma)
    ramps = gi.index_ramps(1, shape)
    out, idx = orc.warp(ramps, flow, mode="nearest", return_idx=True)
    inb = np.ones((1, 1, *shape), bool)
    for d, s in enumerate(shape):
        inb &= (idx[:, d:d + 1] >= 0) & (idx[:, d:d + 1] < s)
    # reference nearest-mode warp of index ramps == the sampled integer index (0 outside)
    assert np.array_equal(inb.astype(np.uint8), g[name + "/nearest_inb"])
    assert np.array_equal(np.where(inb, idx, 0).astype(np.int16), g[name + "/nearest_idx"])
    assert np.array_equal(out.astype(np.int16), g[name + "/nearest_idx"])
    if name + "/ngrid" in g:
        assert np.array_equal(orc.normalized_grid(flow), g[name + "/ngrid"])  # bit-exact fp32


@pytest.mark.parametrize("case", gi.HALF_CASES, ids=[c[0] for c in gi.HALF_CASES])
def test_warp_half_integer_rounding(case, golden, orc):
    name, shape, seed = case
    g = golden("warp")
    flow = gi.half_integer_flow(seed, 1, shape)
    out = orc.warp(gi.index_ramps(1, shape), flow, mode="nearest")
    assert np.array_equal(out.astype(np.int16), g[name + "/nearest_idx"])
    ones = orc.warp(np.ones((1, 1, *shape), np.float32), flow, mode="nearest")
    assert np.array_equal(ones.astype(np.uint8), g[name + "/nearest_inb"])


@pytest.mark.parametrize("case", gi.WARP_CASES, ids=[c[0] for c in gi.WARP_CASES])
def test_warp_linear_values(case, golden, orc):
    name, shape, sigma, seed = case
    g = golden("warp")
    out = orc.warp(gi.image(seed + 100, 1, shape), gi.flow(seed, 1, shape, sigma))
    np.testing.assert_allclose(out, g[name + "/linear"], atol=2e-6, rtol=0)


def test_zero_flow_is_not_identity(orc):
    """normalise/unnormalise is not the identity in fp32 at S=256."""
    img = gi.image(5, 1, (256, 256))
    out, idx = orc.warp(img, np.zeros((1, 2, 256, 256), np.float32), return_idx=True)
    cols = np.arange(256)[None, :].repeat(256, 0)
    assert (idx[0, 1] != cols).sum() > 0          # some floors land on k-1
    assert 0 < np.abs(out - img).max() < 1e-4


def test_integer_shift_is_roll(orc):
    img = gi.image(6, 1, (32, 40))
    flow = np.zeros((1, 2, 32, 40), np.float32)
    flow[:, 0] = 1.0
    flow[:, 1] = -2.0
    out = orc.warp(img, flow)
    ref = np.zeros_like(img)
    ref[:, :, :31, 2:] = img[:, :, 1:, :38]
    np.testing.assert_allclose(out[:, :, :30, 3:], ref[:, :, :30, 3:], atol=1e-5)
    assert np.all(out[:, :, 31, :] == 0) and np.all(out[:, :, :, :2][..., :1] == 0)


@pytest.mark.parametrize("name,shape,sigma,seed", [("v2d", (128, 128), 8.0, 51), ("v3d", (16, 20, 24), 4.0, 52),
                                                  ("v2d_tiny", (128, 128), 1e-3, 53)])
def test_vecint(name, shape, sigma, seed, golden, orc):
    g = golden("vecint_resize")
    vec = gi.smooth_field(gi.rng(seed), (2, len(shape), *shape), sigma)
    # Seven compositions amplify 1-ulp differences in interpolation weights where a sample point
    # sits on the volume border (a corner flips in/out of bounds): allow 0.2% of voxels to exceed
    # 2e-5, none to exceed 1e-3.
    for got, want in ((orc.vecint(vec, 7), g[name + "/out"]), (orc.vecint(-vec, 7), g[name + "/out_neg"])):
        err = np.abs(got - want)
        assert err.max() < 1e-3
        assert (err > 2e-5).mean() < 2e-3


@pytest.mark.parametrize("name,shape,seed", [("r2d", (64, 96), 61), ("r3d", (16, 20, 24), 62), ("r2d_odd", (37, 53), 63)])
def test_resize(name, shape, seed, golden, orc):
    g = golden("vecint_resize")
    x = gi.weights(seed, (2, len(shape), *shape), 1.0)
    down = orc.resize_transform(x, 2)
    up = orc.resize_transform(x, 0.5)
    assert down.shape == g[name + "/down"].shape and up.shape == g[name + "/up"].shape
    np.testing.assert_allclose(down, g[name + "/down"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(up, g[name + "/up"], atol=2e-6, rtol=0)


@pytest.mark.parametrize("name,shape,seed", [("n2d", (64, 64), 71), ("n3d", (24, 28, 32), 72), ("n2d_odd", (45, 70), 73)])
def test_ncc(name, shape, seed, golden, orc):
    g = golden("losses")
    I, J = gi.image_textured(seed, 2, shape), gi.image_textured(seed + 1, 2, shape)
    out, cc = orc.ncc(I, J, return_cc=True)
    assert abs(out[0] - g[name + "/loss"]) <= 1e-4          # the north_star tolerance
    np.testing.assert_allclose(cc, g[name + "/cc"], atol=2e-3, rtol=1e-3)
    mask = (gi.image(seed + 2, 2, shape) > -0.5).astype(np.float32)
    assert abs(orc.ncc(I, J, mask=mask)[0] - g[name + "/loss_masked"]) <= 1e-4
    assert abs(orc.ncc(I, I)[0] - g[name + "/loss_self"]) <= 1e-4


def test_ncc_survey_known_answers(golden, orc):
    g = golden("losses")
    out = orc.ncc(g["survey/ncc_rand_a"], g["survey/ncc_rand_b"])
    assert abs(out[0] - g["survey/ncc_rand"]) <= 1e-5
    assert abs(float(g["survey/ncc_rand"]) - (-0.24840017)) < 1e-6
    gl = orc.grad_loss(g["survey/grad_in"], penalty=2)
    assert abs(gl - float(g["survey/grad_l2_3d"])) <= 1e-5
    assert abs(float(g["survey/grad_l2_3d"]) - 2.0) < 0.05   # E[(a-b)^2] = 2 for unit normals


def test_ncc_constant_image_eps_path(orc):
    I = np.full((1, 1, 32, 32), 0.37, np.float32)
    out, cc = orc.ncc(I, I, return_cc=True)
    assert np.isfinite(out[0]) and -1.0 <= out[0] <= 0.0
    assert np.all(cc[..., 4:-4, 4:-4] < 1e-2)   # interior: I_var ~ 0 -> cc = cross^2/eps ~ 0
    assert np.all(cc[..., 0, :] > 0.9)          # border windows see the zero padding: cc ~ 1


@pytest.mark.parametrize("name,shape,seed", [("g2d", (64, 80), 81), ("g3d", (12, 16, 20), 82)])
def test_grad_loss(name, shape, seed, golden, orc):
    g = golden("losses")
    x = gi.weights(seed, (2, len(shape), *shape), 1.0)
    assert abs(orc.grad_loss(x, 1, 2.0) - g[name + "/l1"]) <= 2e-6 * abs(g[name + "/l1"])
    assert abs(orc.grad_loss(x, 2, 1.0) - g[name + "/l2"]) <= 2e-6 * abs(g[name + "/l2"])


def test_smoothing_and_l1(golden, orc):
    g = golden("losses")
    x = gi.weights(91, (2, 2, 64, 80), 1.0)
    assert abs(orc.grad_loss(x, 2, 1.0) - g["smooth/loss"]) <= 2e-6 * abs(g["smooth/loss"])
    a, b = gi.image(101, 2, (64, 64)), gi.image(102, 2, (64, 64))
    out = orc.l1_masked(a, b, mu=b, mv=a)
    assert out[1] == g["l1/mask_sum"]
    assert abs(out[0] - g["l1/loss"]) <= 1e-6
    mask = (b > -0.95) | (a > -0.95)
    assert abs(orc.l1_masked(a, b, mask=mask)[0] - g["l1/loss"]) <= 1e-6
    assert orc.l1_masked(a, b, mask=np.zeros_like(mask))[0] == 0.0 == g["l1/empty"]
    assert abs(orc.l1_masked(a, b)[0] - g["l1/nomask"]) <= 1e-6
