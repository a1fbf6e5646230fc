"""GPU parity of the network kernels and host mirrors (through the C ABI) against (a) golden vectors
produced by the reference's own modules, (b) the C oracle, and (c) a plain torch fp32 CPU
restatement for the floating-point layer kernels.  fp32 CUDA-core engine unless stated."""
import argparse

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import inputs as gi

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sd_of(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + "/")}


def close(got, want, atol, what=""):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else want
    np.testing.assert_allclose(got, want, atol=atol, rtol=0, err_msg=what)


@pytest.fixture(scope="module")
def Fn():
    import dfmir_b200.functional as Fn
    return Fn


@pytest.fixture(autouse=True)
def fp32_engine(request):
    """Parity against the fp32 goldens runs on the exact-fp32 CUDA-core convolution engine; tests marked
    `tf32` run the tcgen05 tensor-core engine (TF32 operands, the reference's own cuDNN arithmetic class)
    with the wider tolerance stated in the test."""
    import dfmir_b200.functional as Fn
    prev = Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto" if request.node.get_closest_marker("tf32") else "simt"
    yield
    Fn.CONV_ENGINE = prev


CONV_CASES = [  # nd, N, Cin, Cout, spatial, k, stride, pad, act, planar
    (2, 2, 1, 8, (20, 24), 7, 1, 0, 0, False),      # stem: Cin = 1, 7x7
    (2, 2, 8, 1, (20, 24), 7, 1, 0, 2, False),      # head: Cout = 1, tanh
    (2, 2, 16, 40, (18, 22), 3, 1, 1, 0, False),
    (2, 1, 34, 16, (16, 16), 3, 1, 1, 1, False),    # VoxelMorph extras, LeakyReLU
    (2, 2, 2, 16, (32, 32), 3, 2, 1, 1, False),     # stride-2 encoder
    (2, 2, 16, 2, (16, 20), 3, 1, 1, 0, True),      # flow head, planar output
    (2, 1, 72, 80, (9, 11), 3, 1, 1, 0, False),
    (3, 1, 2, 16, (12, 16, 20), 3, 2, 1, 1, False),
    (3, 1, 18, 8, (6, 8, 10), 3, 1, 1, 1, False),
    (3, 2, 8, 3, (6, 8, 10), 3, 1, 1, 0, True),
    (3, 1, 5, 7, (7, 9, 11), 3, 2, 1, 0, False),    # odd sizes, stride 2
    (2, 2, 1, 64, (45, 76), 7, 1, 0, 0, False),     # stem at full width: several (partial) tiles of the direct kernels
    (2, 1, 64, 1, (76, 45), 7, 1, 0, 2, False),     # head at full width
    (2, 1, 1, 24, (40, 40), 7, 1, 3, 0, False),     # zero-padded thin conv
    (2, 1, 34, 16, (64, 64), 3, 1, 1, 1, False),    # few output channels, >= 4096 positions: one-row-per-thread wgrad
    (2, 2, 16, 2, (48, 64), 3, 1, 1, 0, True),      # flow head at size (planar output)
    (3, 1, 18, 3, (16, 16, 20), 3, 1, 1, 0, True),  # 3-D flow head
    (2, 2, 2, 16, (96, 96), 3, 2, 1, 1, False),     # strided first encoder layer
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[f"c{i}" for i in range(len(CONV_CASES))])
def test_conv_fwd_bwd_vs_torch(case, Fn):
    nd, N, Cin, Cout, S, k, stride, pad, act, planar = case
    r = gi.rng(500 + Cin * 7 + Cout)
    x = torch.from_numpy(r.standard_normal((N, Cin, *S)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin) + (k,) * nd) * 0.1).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    conv = F.conv2d if nd == 2 else F.conv3d
    y = conv(x, w, b, stride=stride, padding=pad)
    y = {0: y, 1: F.leaky_relu(y, 0.2), 2: torch.tanh(y)}[act]
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)

    perm_in = (0, *range(2, nd + 2), 1)
    perm_out = (0, nd + 1, *range(1, nd + 1))
    xg = x.detach().cuda().permute(*perm_in).contiguous().requires_grad_()
    wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
    yg = Fn.conv_cl(xg, wg, bg, stride=stride, pad=pad, act=act, planar_out=planar)
    yg_nc = yg if planar else yg.permute(*perm_out)
    scale = float(y.abs().max())
    close(yg_nc, y, 2e-5 * max(1.0, scale), "forward")
    yg_nc.backward(gy.cuda())
    close(xg.grad.permute(*perm_out), x.grad, 2e-5 * max(1.0, float(x.grad.abs().max())), "dgrad")
    close(wg.grad, w.grad, 1e-4 * max(1.0, float(w.grad.abs().max())), "wgrad")
    close(bg.grad, b.grad, 1e-4 * max(1.0, float(b.grad.abs().max())), "bias grad")


def test_conv_strided_views(Fn):
    """conv reads an interior view of a padded buffer (element strides, no copy)."""
    r = gi.rng(77)
    xp = torch.from_numpy(r.standard_normal((2, 14, 18, 8)).astype(np.float32)).cuda()
    w = torch.from_numpy((r.standard_normal((12, 8, 3, 3)) * 0.1).astype(np.float32)).cuda()
    view = xp[:, 1:-1, 1:-1, :]
    y = Fn.conv_cl(view, w, None, pad=1)
    ref = F.conv2d(view.permute(0, 3, 1, 2).cpu(), w.cpu(), padding=1)
    close(y.permute(0, 3, 1, 2), ref, 2e-5)


@pytest.mark.parametrize("C", [12, 16, 64])       # 12: scalar kernels; 16 / 64: the 128-bit kernels
@pytest.mark.parametrize("relu,pad,with_res", [(True, 0, False), (True, 1, False), (True, 3, False), (False, 1, True), (False, 0, True)])
def test_instnorm_fused(relu, pad, with_res, C, Fn):
    r = gi.rng(600 + pad)
    N, H, W = 2, 10, 14
    x = torch.from_numpy((r.standard_normal((N, C, H, W)) * 2 + 0.5).astype(np.float32)).requires_grad_()
    res = torch.from_numpy(r.standard_normal((N, C, H, W)).astype(np.float32)).requires_grad_() if with_res else None
    y = F.instance_norm(x, eps=1e-5)
    if relu:
        y = F.relu(y)
    if with_res:
        y = y + res
    if pad:
        y = F.pad(y, (pad,) * 4, mode="reflect")
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)

    xg = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    rg = None
    if with_res:   # the residual arrives as a buffer carrying a reflected halo of 1
        rg = F.pad(res.detach(), (1,) * 4, mode="reflect").cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
    yg = Fn.instnorm_cl(xg, relu=relu, out_pad=pad, res=rg, res_pad=1 if with_res else 0)
    close(yg.permute(0, 3, 1, 2), y, 2e-5)
    yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
    close(xg.grad.permute(0, 3, 1, 2), x.grad, 5e-5 * max(1.0, float(x.grad.abs().max())))
    if with_res:
        want = F.pad(res.grad, (1,) * 4)   # gradient lands on the interior only
        close(rg.grad.permute(0, 3, 1, 2), want, 1e-6)


def test_pad_blur_layers_vs_reference(golden, orc, Fn):
    g = golden("nets")
    x = gi.weights(201, (2, 6, 12, 16), 1.0)
    for name, fn, gseed in (("down", Fn.blur_down_cl, 202), ("up", Fn.blur_up_cl, 202)):
        xg = cu(x).permute(0, 2, 3, 1).contiguous().requires_grad_()
        y = fn(xg)
        close(y.permute(0, 3, 1, 2), g[f"layer/{name}"], 1e-6, name)
        y.backward(cu(gi.weights(gseed, tuple(g[f"layer/{name}"].shape), 1.0)).permute(0, 2, 3, 1).contiguous())
        close(xg.grad.permute(0, 3, 1, 2), g[f"layer/{name}_dx"], 2e-6, name + " bwd")
    for p in (1, 3):
        xt = torch.from_numpy(x).requires_grad_()
        ref = F.pad(xt, (p,) * 4, mode="reflect")
        gy = torch.from_numpy(gi.weights(204, tuple(ref.shape), 1.0))
        ref.backward(gy)
        xg = cu(x).permute(0, 2, 3, 1).contiguous().requires_grad_()
        y = Fn.pad_reflect_cl(xg, p)
        assert np.array_equal(y.permute(0, 3, 1, 2).detach().cpu().numpy(), orc.pad_reflect(x, p))
        y.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
        close(xg.grad.permute(0, 3, 1, 2), xt.grad, 1e-6)


@pytest.mark.parametrize("shape,C1,C2,pad_to", [((8, 12), 5, 3, 1), ((4, 6, 8), 5, 3, 1), ((8, 12), 32, 2, 4), ((4, 6, 8), 32, 2, 4),
                                                 ((6, 4, 10), 64, 32, 4), ((16, 8), 8, 6, 4)])
def test_upsample_concat(shape, C1, C2, pad_to, Fn):
    """nearest x2 + concat (vxm networks.py:99-102): scalar kernel and the 128-bit kernels (C1 % 4 == 0, channel stride padded
    to a multiple of 4: the zero channels the next convolution's TMA loads expect)."""
    nd = len(shape)
    r = gi.rng(700 + nd + C1)
    a = torch.from_numpy(r.standard_normal((2, C1, *[s // 2 for s in shape])).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal((2, C2, *shape)).astype(np.float32)).requires_grad_()
    y = torch.cat([F.interpolate(a, scale_factor=2, mode="nearest"), b], dim=1)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    pi, po = (0, *range(2, nd + 2), 1), (0, nd + 1, *range(1, nd + 1))
    ag = a.detach().cuda().permute(*pi).contiguous().requires_grad_()
    bg = b.detach().cuda().permute(*pi).contiguous().requires_grad_()
    yg = Fn.upsample_concat_cl(ag, bg, pad_channels_to=pad_to)
    Cs = (C1 + C2 + pad_to - 1) // pad_to * pad_to
    assert yg.shape[-1] == Cs and float(yg[..., C1 + C2:].abs().sum()) == 0.0
    assert torch.equal(yg[..., :C1 + C2].permute(*po).cpu(), y.detach())
    gyp = torch.zeros(yg.shape, device="cuda")
    gyp[..., :C1 + C2] = gy.cuda().permute(*pi)
    gyp[..., C1 + C2:] = 7.0                     # gradient of the padding channels must be ignored
    yg.backward(gyp)
    close(ag.grad.permute(*po), a.grad, 1e-5)
    close(bg.grad.permute(*po), b.grad, 0)


def test_resnet_block_vs_reference(golden):
    from dfmir_b200 import networks
    g = golden("nets")
    blk = networks.ResnetBlock(6, 'reflect', networks.get_norm_layer('instance'), False, True)
    blk.load_state_dict(sd_of(g, "layer/resblock_sd"))
    blk.cuda()
    x = cu(gi.weights(201, (2, 6, 12, 16), 1.0)).requires_grad_()
    y = blk(x)
    close(y, g["layer/resblock"], 2e-5)
    y.backward(cu(gi.weights(203, tuple(y.shape), 1.0)))
    close(x.grad, g["layer/resblock_dx"], 1e-4 * np.abs(g["layer/resblock_dx"]).max())
    for k, p in blk.named_parameters():
        want = g[f"layer/resblock_grad/{k}"]
        # biases feed an InstanceNorm: exactly-zero true gradient, the golden holds rounding noise
        ref_scale = np.abs(g[f"layer/resblock_grad/{k[:-4]}weight"]).max() if k.endswith("bias") else np.abs(want).max()
        close(p.grad, want, 5e-4 * max(1e-3, ref_scale), k)


def test_resnet_generator_vs_reference(golden):
    from dfmir_b200 import networks
    g = golden("nets")
    G = networks.define_G(1, 1, 8, 'resnet_4blocks', 'instance', False, 'xavier', 0.5, False, False, [], None)
    G.load_state_dict(sd_of(g, "G/sd"))
    G.cuda()
    layers = [0, 4, 8, 12, 14]
    x = cu(g["G/in"]).requires_grad_()
    fake, feats = G(x, layers, encode_only=False)
    close(fake, g["G/fake"], 5e-5)
    assert fake.shape == g["G/fake"].shape
    for i, f in enumerate(feats):
        ref = g[f"G/feat{i}"]
        assert tuple(f.shape) == ref.shape
        close(f, ref, 5e-5 * max(1.0, np.abs(ref).max()), f"feature {layers[i]}")
    enc = G(x, layers, encode_only=True)
    assert len(enc) == len(feats) and all(torch.equal(a, b) for a, b in zip(enc, feats))
    plain = G(x)
    assert torch.equal(plain, fake)
    loss = (fake * cu(gi.weights(212, tuple(fake.shape), 1.0))).sum() + sum(
        (f * cu(gi.weights(213 + i, tuple(f.shape), 0.1))).sum() for i, f in enumerate(feats))
    loss.backward()
    close(x.grad, g["G/dx"], 3e-4 * np.abs(g["G/dx"]).max(), "d input")
    for k, p in G.named_parameters():
        want = g[f"G/grad/{k}"]
        ref_scale = np.abs(g[f"G/grad/{k[:-4]}weight"]).max() if k.endswith("bias") else np.abs(want).max()
        close(p.grad, want, 5e-4 * max(1e-4, ref_scale), k)


def test_patch_sample_and_nce_vs_reference(golden, orc, monkeypatch):
    from dfmir_b200 import networks
    from dfmir_b200.patchnce import PatchNCELoss
    g = golden("nets")
    opt = argparse.Namespace(netF_nc=32, batch_size=2, nce_T=0.07, nce_includes_all_negatives_from_minibatch=False)
    netF = networks.define_F(1, 'mlp_sample', 'instance', False, 'xavier', 0.5, False, [], opt)
    fq = [cu(gi.weights(221, (2, 1, 20, 24), 1.0)).requires_grad_(), cu(gi.weights(222, (2, 16, 10, 12), 1.0)).requires_grad_()]
    fk = [cu(gi.weights(223, (2, 1, 20, 24), 1.0)), cu(gi.weights(224, (2, 16, 10, 12), 1.0))]
    netF.create_mlp(fk)
    netF.load_state_dict(sd_of(g, "F/sd"))
    netF.cuda()
    cnt = [0]
    monkeypatch.setattr(torch, "randperm", gi.det_randperm(cnt))
    k_pool, ids = netF(fk, 48, None)
    monkeypatch.undo()
    q_pool, _ = netF(fq, 48, ids)
    crit = PatchNCELoss(opt)
    total = 0
    for i, (q, k) in enumerate(zip(q_pool, k_pool)):
        assert np.array_equal(ids[i].cpu().numpy(), g[f"F/ids{i}"])
        close(q, g[f"F/q{i}"], 5e-6, "q"); close(k, g[f"F/k{i}"], 5e-6, "k")
        l = crit(q, k)
        close(l, g[f"F/loss{i}"], 1e-4, "loss vs reference")
        close(l, orc.patchnce(q.detach().cpu().numpy(), k.detach().cpu().numpy(), 2, 0.07), 1e-4, "loss vs oracle")
        total = total + l.mean()
    total.backward()
    # gradients of a softmax over 48 logits at T = 0.07: fp32 summation order moves single entries by ~4e-3 of the scale
    close(fq[0].grad, g["F/dq0"], 1e-2 * np.abs(g["F/dq0"]).max(), "dq0")
    close(fq[1].grad, g["F/dq1"], 1e-2 * np.abs(g["F/dq1"]).max(), "dq1")
    for k, p in netF.named_parameters():
        want = g[f"F/grad/{k}"]
        close(p.grad, want, 1e-3 * max(1e-6, np.abs(want).max()), k)


def test_vxm_dense_3d_default_features_vs_reference(golden):
    """VxmDense with the reference's default features (the BASELINE configs[3] network) on 16 x 32 x 16, pinned to the
    reference's own forward / backward (oracle/gen_golden.py: nets3d)."""
    from dfmir_b200 import vxm
    g = golden("nets3d")
    shape = (16, 32, 16)
    R = vxm.VxmDense(shape, int_steps=7, bidir=False)
    missing = R.load_state_dict(sd_of(g, "R3d/sd"), strict=False)
    assert all(k.endswith(".grid") for k in missing.missing_keys) and not missing.unexpected_keys
    R.cuda()
    src = cu(gi.image_textured(241, 1, shape)).requires_grad_()
    tgt = cu(gi.image_textured(242, 1, shape))
    ys, flow = R(src, tgt, registration=True)
    close(flow, g["R3d/pos_flow"], 2e-5, "pos_flow")
    close(ys, g["R3d/y_source"], 2e-5)
    ys_f, flow_f, _, _ = R.forward_with_losses(src, tgt, win=9)          # the fused launch gives the same field / volume
    assert torch.equal(flow_f, flow) and torch.equal(ys_f, ys)
    loss = (ys * cu(gi.weights(243, tuple(ys.shape), 1.0))).sum() + (flow * cu(gi.weights(245, tuple(flow.shape), 0.1))).sum()
    loss.backward()
    close(src.grad, g["R3d/d_src"], 1e-4 * np.abs(g["R3d/d_src"]).max())
    for k, p in R.named_parameters():
        want = g[f"R3d/grad/{k}"]
        close(p.grad, want, 1e-3 * max(1e-6, np.abs(want).max()), k)


@pytest.mark.parametrize("name,shape,feats", [("R2", (64, 64), [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]),
                                              ("R3", (16, 16, 16), [[8, 16, 16], [16, 16, 16, 8, 8]])])
def test_vxm_dense_vs_reference(name, shape, feats, golden):
    from dfmir_b200 import vxm
    g = golden("nets")
    R = vxm.VxmDense(shape, feats, int_steps=7, bidir=True)
    missing = R.load_state_dict(sd_of(g, name + "/sd"), strict=False)
    assert all(k.endswith(".grid") for k in missing.missing_keys) and not missing.unexpected_keys
    R.cuda()
    src = cu(gi.image_textured(231, 2, shape)).requires_grad_()
    tgt = cu(gi.image_textured(232, 2, shape))
    ys, yt, flow = R(src, tgt)
    close(flow, g[name + "/pos_flow"], 2e-5, "pos_flow")
    close(ys, g[name + "/y_source"], 2e-5); close(yt, g[name + "/y_target"], 2e-5)
    ys2, flow2 = R(src, tgt, registration=True)
    assert torch.equal(flow2, flow)
    loss = (ys * cu(gi.weights(233, tuple(ys.shape), 1.0))).sum() + (yt * cu(gi.weights(234, tuple(yt.shape), 1.0))).sum() \
        + (flow * cu(gi.weights(235, tuple(flow.shape), 0.1))).sum()
    loss.backward()
    close(src.grad, g[name + "/d_src"], 1e-4 * np.abs(g[name + "/d_src"]).max())
    for k, p in R.named_parameters():
        want = g[f"{name}/grad/{k}"]
        close(p.grad, want, 1e-3 * max(1e-6, np.abs(want).max()), k)


def test_training_step_vs_reference(golden, monkeypatch):
    """One full optimize_parameters (B = 2, 64x64, ngf 8) against the reference's own step on CPU:
    the six logged losses, the visuals, every parameter gradient and the parameters after Adam."""
    from dfmir_b200 import registration_model as rm
    g = golden("step")
    B, S = 2, 64
    opt = rm.default_options(batch_size=B, ngf=8, crop_size=S, load_size=S, netF_nc=32, num_patches=64, gpu_ids=[0])
    dvf_img = torch.from_numpy(g["dvf_img"])
    monkeypatch.setattr(rm, "open_image_to_torch", lambda path, size: dvf_img)
    cnt = [0]
    monkeypatch.setattr(torch, "randperm", gi.det_randperm(cnt))
    m = rm.REGISTRATIONModel(opt)
    data = {'A': torch.from_numpy(gi.image_textured(302, B, (S, S))), 'B': torch.from_numpy(gi.image_textured(303, B, (S, S)))}
    m.data_dependent_initialize(data)
    for n in ('G', 'F', 'R'):
        res = getattr(m, 'net' + n).load_state_dict(sd_of(g, f"sd0/{n}"), strict=False)
        assert all(k.endswith(".grid") for k in res.missing_keys) and not res.unexpected_keys
    m.setup(opt)
    cnt[0] = 100
    m.set_input(data)
    m.optimize_parameters()
    losses = m.get_current_losses()
    for k, v in losses.items():
        assert abs(v - float(g[f"loss/{k}"])) <= 2e-4 * max(1.0, abs(float(g[f"loss/{k}"]))), (k, v, float(g[f"loss/{k}"]))
    for n in ('fake_B', 'idt_B', 'registered', 'regA', 'dvf'):
        close(getattr(m, n), g[f"vis/{n}"], 1e-4, n)
    worst = 0.0
    for n in ('G', 'F', 'R'):
        for k, p in getattr(m, 'net' + n).named_parameters():
            want = g[f"grad/{n}/{k}"]
            tol, exact_zero = gi.grad_tolerance(g, n, k, 5e-3)
            err = float(np.abs(p.grad.cpu().numpy() - want).max())
            worst = max(worst, err / tol)
            assert err <= tol, (n, k, err, tol)
            if not exact_zero:
                # Adam's first step moves every weight by ~lr * sign(g); compare where |g| is not negligible
                after = g[f"sd1/{n}/{k}"]
                sel = np.abs(want) > 1e-2 * np.abs(want).max()
                np.testing.assert_allclose(p.detach().cpu().numpy()[sel], after[sel], atol=2e-5, rtol=0, err_msg=f"{n}.{k}")
    print("worst gradient error / tolerance", worst)


@pytest.mark.tf32
def test_training_step_tensor_core_engine(golden, monkeypatch):
    """The same step on the tcgen05 engine (TF32 operands, fp32 accumulate: what cuDNN does for the
    reference on a GPU).  Tolerances: logged losses within 2e-3 relative, visuals within 5e-3, every
    weight gradient within 0.5 of its tensor's scale with cosine >= 0.97 to the fp32 gradient (TF32 keeps
    10 mantissa bits per operand, the error compounds through 9 residual blocks and their instance norms,
    and at random initialisation the weight-gradient sums cancel ~100x; tests/test_gpu_umma.py bounds the
    same quantity by the error of a TF32-truncated CPU run, and checks each kernel to 3e-5)."""
    from dfmir_b200 import registration_model as rm
    import dfmir_b200.functional as Fn
    g = golden("step")
    B, S = 2, 64
    opt = rm.default_options(batch_size=B, ngf=8, crop_size=S, load_size=S, netF_nc=32, num_patches=64, gpu_ids=[0])
    dvf_img = torch.from_numpy(g["dvf_img"])
    monkeypatch.setattr(rm, "open_image_to_torch", lambda path, size: dvf_img)
    cnt = [0]
    monkeypatch.setattr(torch, "randperm", gi.det_randperm(cnt))
    Fn.UMMA_MIN_POSITIONS = 0
    m = rm.REGISTRATIONModel(opt)
    data = {'A': torch.from_numpy(gi.image_textured(302, B, (S, S))), 'B': torch.from_numpy(gi.image_textured(303, B, (S, S)))}
    m.data_dependent_initialize(data)
    for n in ('G', 'F', 'R'):
        getattr(m, 'net' + n).load_state_dict(sd_of(g, f"sd0/{n}"), strict=False)
    m.setup(opt)
    cnt[0] = 100
    m.set_input(data)
    prof = Fn.ConvProfile()
    Fn.PROFILE = prof
    try:
        m.optimize_parameters()
    finally:
        Fn.PROFILE = None
        Fn.UMMA_MIN_POSITIONS = 4096
    # ngf = 8 keeps G off the tensor-core tile sizes; VoxelMorph's 64-channel layers use them
    assert prof.umma_calls > 0, "no convolution ran on the tcgen05 engine"
    losses = m.get_current_losses()
    for k, v in losses.items():
        ref = float(g[f"loss/{k}"])
        assert abs(v - ref) <= 2e-3 * max(1.0, abs(ref)), (k, v, ref)
    for n in ('fake_B', 'idt_B', 'registered', 'regA'):
        close(getattr(m, n), g[f"vis/{n}"], 5e-3, n)
    for n in ('G', 'F', 'R'):
        for k, p in getattr(m, 'net' + n).named_parameters():
            want = g[f"grad/{n}/{k}"]
            tol, exact_zero = gi.grad_tolerance(g, n, k, 0.5)
            if exact_zero or (n == 'G' and k.endswith('.bias')):
                # biases in front of an InstanceNorm: true gradient 0, both sides hold rounding noise; the head conv's
                # bias: a sum over every pixel that cancels ~100x (inputs.grad_tolerance), meaningless at TF32
                continue
            got = p.grad.cpu().numpy()
            assert float(np.abs(got - want).max()) <= tol, (n, k)
            if not exact_zero and k.endswith("weight") and np.abs(want).max() > 1e-6:
                cos = float((got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want) + 1e-30))
                assert cos >= 0.97, (n, k, cos)


@pytest.mark.parametrize("C,H,W", [(8, 12, 16), (16, 9, 11)])
def test_blur_layers_128bit_path(C, H, W, Fn):
    """Downsample / Upsample (anti-aliased) on the float4 kernels (C % 4 == 0), odd and even sizes, against the
    torch CPU port of the reference layers (oracle/torch_port.py blur_down / blur_up) with autograd."""
    from oracle import torch_port as tp
    r = gi.rng(800 + C)
    for name, ref_fn, fn in (("down", tp.blur_down, Fn.blur_down_cl), ("up", tp.blur_up, Fn.blur_up_cl)):
        x = torch.from_numpy(r.standard_normal((2, C, H, W)).astype(np.float32)).requires_grad_()
        y = ref_fn(x)
        gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
        y.backward(gy)
        xg = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        yg = fn(xg)
        close(yg.permute(0, 3, 1, 2), y, 2e-6, name)
        yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
        close(xg.grad.permute(0, 3, 1, 2), x.grad, 4e-6, name + " bwd")
