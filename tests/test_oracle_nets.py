"""CPU: the oracle's restatement of the networks (oracle/nets_oracle.py over the C layers) against
golden vectors produced by the reference's own modules (oracle/gen_golden.py nets)."""
import numpy as np
import pytest

import inputs as gi


@pytest.fixture(scope="module")
def nets(orc):
    from oracle import nets_oracle
    return nets_oracle


def sd_of(g, prefix):
    return {k[len(prefix) + 1:]: g[k] for k in g.files if k.startswith(prefix + "/")}


def test_blur_layers(golden, orc):
    g = golden("nets")
    x = gi.weights(201, (2, 6, 12, 16), 1.0)
    np.testing.assert_allclose(orc.blur_down(x), g["layer/down"], atol=1e-6)
    np.testing.assert_allclose(orc.blur_up(x), g["layer/up"], atol=1e-6)


def test_resnet_block(golden, orc):
    g = golden("nets")
    sd = sd_of(g, "layer/resblock_sd")
    x = gi.weights(201, (2, 6, 12, 16), 1.0)
    h = orc.conv(orc.pad_reflect(x, 1), sd["conv_block.1.weight"], sd["conv_block.1.bias"])
    h = np.maximum(orc.instnorm(h), 0)
    h = orc.conv(orc.pad_reflect(h, 1), sd["conv_block.5.weight"], sd["conv_block.5.bias"])
    np.testing.assert_allclose(x + orc.instnorm(h), g["layer/resblock"], atol=2e-5)


def test_resnet_generator(golden, nets):
    g = golden("nets")
    sd = sd_of(g, "G/sd")
    layers = [0, 4, 8, 12, 14]
    fake, feats = nets.resnet_generator(g["G/in"], sd, n_blocks=4, layers=layers)
    np.testing.assert_allclose(fake, g["G/fake"], atol=5e-5)
    for i, l in enumerate(layers):
        ref = g[f"G/feat{i}"]
        np.testing.assert_allclose(feats[l], ref, atol=5e-5 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize("name,shape,n_enc,n_dec", [("R2", (64, 64), 6, 7), ("R3", (16, 16, 16), 3, 5)])
def test_vxm_dense(name, shape, n_enc, n_dec, golden, nets):
    g = golden("nets")
    sd = sd_of(g, name + "/sd")
    src, tgt = gi.image_textured(231, 2, shape), gi.image_textured(232, 2, shape)
    ys, yt, flow = nets.vxm_dense(src, tgt, sd, n_enc, n_dec)
    np.testing.assert_allclose(flow, g[name + "/pos_flow"], atol=2e-5)
    np.testing.assert_allclose(ys, g[name + "/y_source"], atol=2e-5)
    np.testing.assert_allclose(yt, g[name + "/y_target"], atol=2e-5)


def test_vxm_dense_3d_default_features(golden, nets):
    """The network of BASELINE configs[3]: the reference's default U-Net features (4 encoder / 7 decoder convs)."""
    g = golden("nets3d")
    sd = sd_of(g, "R3d/sd")
    shape = (16, 32, 16)
    src, tgt = gi.image_textured(241, 1, shape), gi.image_textured(242, 1, shape)
    ys, _, flow = nets.vxm_dense(src, tgt, sd, 4, 7)
    np.testing.assert_allclose(flow, g["R3d/pos_flow"], atol=2e-5)
    np.testing.assert_allclose(ys, g["R3d/y_source"], atol=2e-5)


def test_patch_sample_and_nce(golden, orc, nets):
    g = golden("nets")
    sd = sd_of(g, "F/sd")
    fq = [gi.weights(221, (2, 1, 20, 24), 1.0), gi.weights(222, (2, 16, 10, 12), 1.0)]
    fk = [gi.weights(223, (2, 1, 20, 24), 1.0), gi.weights(224, (2, 16, 10, 12), 1.0)]
    for i in range(2):
        ids = g[f"F/ids{i}"]
        # the reference's stand-in draw (gen_golden.det_randperm): call i+1 -> RandomState(9001 + i)
        assert np.array_equal(ids, np.random.RandomState(9001 + i).permutation(fq[i].shape[2] * fq[i].shape[3])[:48])
        q, k = nets.patch_sample(fq[i], ids, sd, i), nets.patch_sample(fk[i], ids, sd, i)
        np.testing.assert_allclose(q, g[f"F/q{i}"], atol=2e-6)
        np.testing.assert_allclose(k, g[f"F/k{i}"], atol=2e-6)
        np.testing.assert_allclose(orc.patchnce(q, k, 2, 0.07), g[f"F/loss{i}"], atol=2e-5)


def test_torch_port_step(golden, monkeypatch):
    """oracle/torch_port.py (the CPU baseline / autograd checker) reproduces the reference's own
    optimize_parameters: losses, visuals, every gradient, parameters after the three Adam steps."""
    import torch
    from oracle import torch_port as tp
    g = golden("step")
    torch.set_num_threads(4)
    sds = {n: {k[len(f"sd0/{n}/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"sd0/{n}/")} for n in "GFR"}
    st = tp.Step(sds["G"], sds["F"], sds["R"], n_blocks=9, batch_size=2, num_patches=64,
                 dvf_image=torch.from_numpy(g["dvf_img"]))
    cnt = [100]
    monkeypatch.setattr(torch, "randperm", gi.det_randperm(cnt))
    A, B = torch.from_numpy(gi.image_textured(302, 2, (64, 64))), torch.from_numpy(gi.image_textured(303, 2, (64, 64)))
    losses = st.step(A, B)
    for k, v in losses.items():
        assert abs(v - float(g[f"loss/{k}"])) <= 1e-5, (k, v, float(g[f"loss/{k}"]))
    for n, v in st.visuals.items():
        np.testing.assert_allclose(v.detach().numpy(), g[f"vis/{n}"], atol=1e-5)
    for n in "GFR":
        for k, p in st.P[n].items():
            if p.requires_grad:
                want = g[f"grad/{n}/{k}"]
                tol, exact_zero = gi.grad_tolerance(g, n, k, 1e-4)
                np.testing.assert_allclose(p.grad.numpy(), want, atol=tol, err_msg=f"{n}.{k}")
                if not exact_zero:   # Adam turns a noise gradient into a +-lr step: nothing to compare there
                    sel = np.abs(want) > 1e-3 * np.abs(want).max()
                    np.testing.assert_allclose(p.detach().numpy()[sel], g[f"sd1/{n}/{k}"][sel], atol=2e-5, err_msg=f"{n}.{k}")


@pytest.mark.parametrize("name,shape,seed", [("n2d", (64, 64), 71), ("n3d", (24, 28, 32), 72), ("n2d_odd", (45, 70), 73)])
def test_torch_port_ncc_loss(name, shape, seed, golden):
    """oracle/torch_port.ncc_loss (the CPU-baseline composition of bench.py's 3-D legs) against the reference's own
    NCC_Loss values and gradients (tests/golden/losses.npz, generated by oracle/gen_golden.py)."""
    import torch
    from oracle import torch_port as tp
    g = golden("losses")
    I = torch.from_numpy(gi.image_textured(seed, 2, shape)).requires_grad_()
    J = torch.from_numpy(gi.image_textured(seed + 1, 2, shape)).requires_grad_()
    loss = tp.ncc_loss(I, J, 9)
    loss.backward()
    assert abs(float(loss) - float(g[name + "/loss"])) <= 1e-6
    np.testing.assert_allclose(I.grad.numpy(), g[name + "/dI"], atol=1e-6 * max(1.0, np.abs(g[name + "/dI"]).max()), rtol=1e-4)
    np.testing.assert_allclose(J.grad.numpy(), g[name + "/dJ"], atol=1e-6 * max(1.0, np.abs(g[name + "/dJ"]).max()), rtol=1e-4)
    mask = torch.from_numpy((gi.image(seed + 2, 2, shape) > -0.5).astype(np.float32))
    assert abs(float(tp.ncc_loss(I.detach(), J.detach(), 9, mask=mask)) - float(g[name + "/loss_masked"])) <= 1e-6


def test_torch_port_grad_loss(golden):
    import torch
    from oracle import torch_port as tp
    g = golden("losses")
    assert abs(float(tp.grad_loss(torch.from_numpy(g["survey/grad_in"]))) - float(g["survey/grad_l2_3d"])) <= 1e-6
    for name, shape, seed in [("g2d", (64, 80), 81), ("g3d", (12, 16, 20), 82)]:
        x = torch.from_numpy(gi.weights(seed, (2, len(shape), *shape), 1.0)).requires_grad_()
        for pen in ("l1", "l2"):
            x.grad = None
            loss = tp.grad_loss(x, pen) * (2.0 if pen == "l1" else 1.0)     # the l1 goldens carry loss_mult = 2
            loss.backward()
            assert abs(float(loss) - float(g[f"{name}/{pen}"])) <= 1e-6 * max(1.0, abs(float(g[f"{name}/{pen}"])))
            np.testing.assert_allclose(x.grad.numpy(), g[f"{name}/{pen}_dx"], atol=1e-9, rtol=1e-5)


def test_torch_port_step3d_matches_reference_vxm(golden):
    """Step3D.forward (the 3-D CPU baseline of bench.py) against the reference's own VxmDense-3D forward / backward
    with default features at 16 x 32 x 16 (tests/golden/nets3d.npz, oracle/gen_golden.py gen_nets3d)."""
    import torch
    from oracle import torch_port as tp
    g = golden("nets3d")
    sd = {k[len("R3d/sd/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("R3d/sd/")}
    shape = (16, 32, 16)
    st = tp.Step3D(sd, [[16, 32, 32, 32], [32, 32, 32, 32, 32, 16, 16]])
    src = torch.from_numpy(gi.image_textured(241, 1, shape)).requires_grad_()
    tgt = torch.from_numpy(gi.image_textured(242, 1, shape))
    ys, flow = st.forward(src, tgt)
    np.testing.assert_allclose(ys.detach().numpy(), g["R3d/y_source"], atol=2e-6)
    np.testing.assert_allclose(flow.detach().numpy(), g["R3d/pos_flow"], atol=2e-6)
    loss = (ys * torch.from_numpy(gi.weights(243, tuple(ys.shape), 1.0))).sum() + (flow * torch.from_numpy(gi.weights(245, tuple(flow.shape), 0.1))).sum()
    loss.backward()
    np.testing.assert_allclose(src.grad.numpy(), g["R3d/d_src"], atol=1e-5 * np.abs(g["R3d/d_src"]).max())
    for k, v in st.P.items():
        if v.grad is not None:
            want = g["R3d/grad/" + k]
            np.testing.assert_allclose(v.grad.numpy(), want, atol=1e-4 * max(1e-6, np.abs(want).max()), err_msg=k)
    # and one optimiser step of the composition bench.py times runs
    st.step(src.detach(), tgt)
    assert np.isfinite(st.losses['ncc']) and np.isfinite(st.losses['grad'])
