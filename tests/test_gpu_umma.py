"""GPU parity of the tcgen05 (tensor-core, TF32 operands / fp32 accumulate) convolution engine against
torch's fp32 CPU convolution and against the exact-fp32 CUDA-core engine.  TF32 keeps 10 mantissa bits
of each operand, so the tolerance is relative: 3e-3 of the output scale (the reference's own CUDA path,
cuDNN with allow_tf32, has the same error class)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import inputs as gi

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def small_shapes_on_tensor_cores():
    """The engine keeps layers with < 4096 output positions on the fp32 kernels; the unit shapes here are smaller."""
    import dfmir_b200.functional as Fn
    prev = Fn.UMMA_MIN_POSITIONS
    Fn.UMMA_MIN_POSITIONS = 0
    yield
    Fn.UMMA_MIN_POSITIONS = prev

CASES = [  # N, Cin, Cout, H, W, k, pad   (H, W = input spatial size)
    (2, 64, 128, 16, 128, 3, 1),     # TW = 128, TH = 1
    (2, 128, 256, 8, 64, 3, 1),      # TW = 64, TH = 2
    (1, 256, 256, 18, 18, 3, 0),     # ResnetBlock conv on a reflect-padded buffer: 16x16 output, TW = 16
    (2, 256, 128, 12, 40, 3, 1),     # partial tiles in w
    (1, 128, 64, 9, 256, 3, 1),      # two tiles per row
    (3, 64, 64, 10, 10, 1, 0),       # 1x1
    (1, 256, 256, 66, 66, 3, 0),     # full-size ResnetBlock geometry: dgrad output 66x66 (partial tiles)
    (2, 36, 16, 64, 64, 3, 1),       # VoxelMorph extras.0 on the channel-padded concat: 2 chunks (32 + 4), N tile 16
    (2, 48, 32, 64, 64, 3, 1),       # uparm.5: partial second chunk, N tile 32
    (2, 16, 2, 64, 64, 3, 1),        # flow head: half a chunk of K, two output channels in a 16-wide tile
    (1, 96, 32, 64, 72, 3, 1),       # three chunks, partial tiles in w
    (1, 64, 512, 12, 20, 3, 1),      # CTA-pair kernel with two 256-channel N tiles, partial tiles in h and w
    (3, 32, 256, 10, 12, 3, 1),      # CTA pairs across samples (2 tiles per image), one K chunk
]


@pytest.mark.parametrize("case", CASES, ids=[f"u{i}" for i in range(len(CASES))])
def test_umma_conv_fwd_dgrad(case):
    """Forward, data gradient and weight gradient of one layer on the tcgen05 engine, against
    (a) the exact fp32 result (TF32 tolerance, 3e-3 of the output scale) and
    (b) a float64 CPU convolution of the operands TRUNCATED to TF32 (oracle.torch_port.tf32_round): the
        tensor core keeps the top 10 mantissa bits of each fp32 operand and accumulates in fp32, so this
        must agree to accumulation-order noise (3e-5 of the scale) — the tight check of tile addressing,
        swizzled layouts, tap shifts, zero fill and split-K."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    N, Cin, Cout, H, W, k, pad = case
    r = gi.rng(900 + Cin + Cout + H)
    x = torch.from_numpy(r.standard_normal((N, Cin, H, W)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    y = F.conv2d(x, w, b, padding=pad)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    xq, wq, gq = (tp.tf32_round(t.detach()).double() for t in (x, w, gy))
    emu = (F.conv2d(xq, wq, b.detach().double(), padding=pad),
           torch.nn.grad.conv2d_input(x.shape, wq, gq, padding=pad),
           torch.nn.grad.conv2d_weight(xq, w.shape, gq, padding=pad))

    def run(engine):
        Fn.CONV_ENGINE = engine
        prof = Fn.ConvProfile()
        Fn.PROFILE = prof
        xg = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        yg = Fn.conv_cl(xg, wg, bg, pad=pad)
        yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
        torch.cuda.synchronize()
        Fn.PROFILE = None
        return (yg.detach().permute(0, 3, 1, 2).cpu(), xg.grad.permute(0, 3, 1, 2).cpu(), wg.grad.cpu(), bg.grad.cpu()), prof

    try:
        exact, _ = run("simt")
        tc, prof = run("auto")
    finally:
        Fn.CONV_ENGINE, Fn.PROFILE = "auto", None
    wgrad_tc = bool(_wgrad_supported(Cin, Cout))
    dgrad_tc = Cout % 4 == 0 and Cout >= 16          # the data gradient reduces over Cout
    assert prof.umma_calls == 1 + int(dgrad_tc) + int(wgrad_tc), "forward (and dgrad / wgrad where covered) must run on the tensor-core engine"
    for name, got, want, ex, em in (("fwd", tc[0], y.detach(), exact[0], emu[0]), ("dgrad", tc[1], x.grad, exact[1], emu[1]),
                                    ("wgrad", tc[2], w.grad, exact[2], emu[2])):
        scale = float(want.abs().max())
        assert float((ex - want).abs().max()) <= 2e-4 * scale, name + " (fp32 engine)"
        if (name == "wgrad" and not wgrad_tc) or (name == "dgrad" and not dgrad_tc):   # stays on the fp32 kernel
            assert float((got - want).abs().max()) <= 2e-4 * scale
            continue
        err = float((got - want).abs().max())
        assert err <= 3e-3 * scale, (name, err, scale)
        assert err > 1e-6 * scale, "tensor-core result equals the fp32 one: engine did not run?"
        err_emu = float((got.double() - em).abs().max())
        assert err_emu <= 3e-5 * scale, (name, "vs TF32-truncated float64 reference", err_emu, scale)
    np.testing.assert_allclose(tc[3].numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


def _wgrad_supported(Cin, Cout):
    return Cin % 4 == 0 and Cout % 4 == 0 and Cin >= 16 and Cout >= 16


def test_umma_supported_shapes():
    import ctypes
    import dfmir_b200.functional as Fn
    from dfmir_b200 import umma

    def desc(nd, Cin, Cout, k, stride):
        S = [16] * nd
        O = [(s + 2 * (k // 2) - k) // stride + 1 for s in S]
        xs = [int(np.prod(S)) * Cin] + [int(np.prod(S[i + 1:])) * Cin for i in range(nd)] + [1]
        ys = [int(np.prod(O)) * Cout] + [int(np.prod(O[i + 1:])) * Cout for i in range(nd)] + [1]
        return Fn._make_desc(nd, 1, Cin, Cout, S, O, [k] * nd, [k // 2] * nd, stride, 0, xs, ys)

    assert umma.supported(desc(2, 64, 128, 3, 1)) and umma.supported(desc(2, 64, 128, 3, 1), dgrad=True)
    assert umma.supported(desc(3, 36, 16, 3, 1)) and umma.supported(desc(3, 16, 3, 3, 1))
    assert not umma.supported(desc(3, 16, 3, 3, 1), dgrad=True)        # reduction over 3 channels: rows not 16-byte multiples
    assert not umma.supported(desc(2, 1, 64, 7, 1))                    # stem: one input channel
    assert not umma.supported(desc(2, 34, 16, 3, 1))                   # 34 channels: pixel stride not a 16-byte multiple
    assert not umma.supported(desc(2, 64, 128, 3, 2))                  # strided
    assert not umma.supported(desc(3, 2, 16, 3, 1))                    # first encoder layer


def test_generator_ngf64_tensor_core_vs_cpu_port():
    """ResnetGenerator at the reference width (ngf = 64: all five conv shapes on the tcgen05 engine, fwd +
    dgrad + wgrad) against the float64 CPU port of the reference module (oracle/torch_port.py).
    Truncating operands to TF32 makes the network discontinuous, so two TF32 evaluations only agree in
    error CLASS: at random initialisation the weight-gradient sums cancel ~100x and a TF32 pipeline is
    4-15 % (of each tensor's largest entry) away from float64.  The bar: output within 5e-3, and every
    weight gradient no further from float64 than 4x the CPU port run with TF32-truncated operands."""
    from oracle import torch_port as tp
    from dfmir_b200 import networks
    import dfmir_b200.functional as Fn
    sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=4, crop=64, seed=3)
    x = torch.from_numpy(gi.image_textured(411, 2, (64, 64)))
    wts = torch.from_numpy(gi.weights(412, (2, 1, 64, 64), 1.0))

    def cpu(dtype, mode):
        tp.TF32_EMULATION = mode
        try:
            leaves = {k: v.clone().to(dtype).requires_grad_() for k, v in sdG.items()}
            out = tp.resnet_generator(x.to(dtype), leaves, 4)
            (out * wts.to(dtype)).sum().backward()
        finally:
            tp.TF32_EMULATION = None
        return out.detach().double(), {k: v.grad.double() for k, v in leaves.items()}

    o64, g64 = cpu(torch.float64, None)
    oem, gem = cpu(torch.float32, "trunc")
    G = networks.define_G(1, 1, 64, 'resnet_4blocks', 'instance', False, 'xavier', 0.02, False, False, [], None)
    missing = G.load_state_dict(sdG, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("filt") for k in missing.missing_keys)
    G.cuda()
    prof = Fn.ConvProfile()
    Fn.PROFILE, prev = prof, Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto"
    try:
        out = G(x.cuda())
        (out * wts.cuda()).sum().backward()
    finally:
        Fn.PROFILE, Fn.CONV_ENGINE = None, prev
    assert prof.umma_calls >= 3 * 12, prof.umma_calls      # 12 tensor-core layers x (fwd, dgrad, wgrad)
    scale = float(o64.abs().max())
    assert float((out.detach().cpu().double() - o64).abs().max()) <= 5e-3 * scale
    for k, p in G.named_parameters():
        if not k.endswith("weight"):
            continue
        sc = float(g64[k].abs().max())
        e_tc = float((p.grad.cpu().double() - g64[k]).abs().max()) / sc
        e_emu = float((gem[k] - g64[k]).abs().max()) / sc
        assert e_tc <= 4.0 * e_emu + 2e-3, (k, e_tc, e_emu)
        cos = float((p.grad.cpu().double() * g64[k]).sum() / (p.grad.cpu().double().norm() * g64[k].norm()))
        assert cos >= 0.98, (k, cos)


CASES3D = [  # N, Cin, Cout, D, H, W
    (1, 36, 16, 16, 16, 32),         # VoxelMorph-3D extras.0 geometry (scaled down): 27 taps x 2 chunks
    (2, 48, 32, 8, 16, 16),
    (1, 16, 3, 12, 20, 24),          # 3-D flow head, partial tiles in every axis
    (1, 64, 32, 16, 16, 16),
    # weight gradient through the halo kernel (all taps per CTA, W >= 24): partial tiles in h and w, 1-3 channel chunks
    (1, 36, 32, 6, 10, 40),
    (1, 16, 16, 5, 7, 70),
    (2, 48, 32, 4, 9, 33),
    (1, 32, 64, 4, 8, 32),           # two output-channel tiles
    (1, 96, 16, 3, 6, 24),
    # depth-march kernel (csrc/conv_umma.cu: conv_umma_dmarch_kernel; >= 8192 voxels, depth >= 8): 16 / 32 / 64-channel tiles,
    # 1 - 2 channel chunks, partial (h, w) tiles, several depth chunks, enough steps to wrap the accumulator ring
    (1, 36, 16, 40, 20, 24),
    (2, 16, 36, 12, 16, 40),
    (1, 48, 32, 20, 18, 20),
    (1, 16, 3, 24, 20, 24),
    (1, 64, 64, 44, 16, 16),
]


@pytest.mark.parametrize("case", CASES3D, ids=[f"v{i}" for i in range(len(CASES3D))])
def test_umma_conv3d_fwd_dgrad(case):
    """3-D 3x3x3 stride-1 convolutions on the tcgen05 engine (5-D TMA boxes, 27 taps): forward and data gradient
    against the float64 CPU convolution of TF32-truncated operands (3e-5) and the exact result (3e-3); the weight
    gradient likewise where both channel counts are multiples of 4 (else the exact fp32 kernel)."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    N, Cin, Cout, D, H, W = case
    r = gi.rng(950 + Cin + Cout + D)
    x = torch.from_numpy(r.standard_normal((N, Cin, D, H, W)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, 3, 3, 3)) / np.sqrt(Cin * 27)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    y = F.conv3d(x, w, b, padding=1)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    xq, wq, gq = (tp.tf32_round(t.detach()).double() for t in (x, w, gy))
    emu_y = F.conv3d(xq, wq, b.detach().double(), padding=1)
    emu_dx = torch.nn.grad.conv3d_input(x.shape, wq, gq, padding=1)
    prev = Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto"
    prof = Fn.ConvProfile(); Fn.PROFILE = prof
    try:
        xg = x.detach().cuda().permute(0, 2, 3, 4, 1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        yg = Fn.conv_cl(xg, wg, bg, pad=1)
        yg.backward(gy.cuda().permute(0, 2, 3, 4, 1).contiguous())
        torch.cuda.synchronize()
    finally:
        Fn.CONV_ENGINE, Fn.PROFILE = prev, None
    kinds = prof.by_kind()
    assert "umma_fwd" in kinds, kinds.keys()
    if Cout % 4 == 0 and Cout >= 16:
        assert "umma_dgrad" in kinds, kinds.keys()
    got_y = yg.detach().permute(0, 4, 1, 2, 3).cpu()
    got_dx = xg.grad.permute(0, 4, 1, 2, 3).cpu()
    for name, got, want, em in (("fwd", got_y, y.detach(), emu_y), ("dgrad", got_dx, x.grad, emu_dx)):
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 3e-3 * scale, name
        if name == "fwd" or "umma_dgrad" in kinds:
            assert float((got.double() - em).abs().max()) <= 3e-5 * scale, (name, "vs TF32-truncated float64 reference")
    wscale = float(w.grad.abs().max())
    if _wgrad_supported(Cin, Cout):      # tensor-core weight gradient (M padded to 128 channels, N to 32/64 by zero fill)
        assert "umma_wgrad" in kinds, kinds.keys()
        emu_dw = torch.nn.grad.conv3d_weight(xq, w.shape, gq, padding=1)
        assert float((wg.grad.cpu().double() - emu_dw).abs().max()) <= 3e-5 * wscale
        np.testing.assert_allclose(wg.grad.cpu().numpy(), w.grad.numpy(), atol=3e-3 * wscale)
    else:
        np.testing.assert_allclose(wg.grad.cpu().numpy(), w.grad.numpy(), atol=2e-4 * wscale)
    np.testing.assert_allclose(bg.grad.cpu().numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


@pytest.mark.parametrize("shape,Cin,Cout", [((64, 64), 16, 2), ((12, 20, 40), 16, 3), ((6, 10, 72), 32, 3)])
def test_planar_flow_head_backward_on_tensor_cores(shape, Cin, Cout):
    """The VoxelMorph flow head (vxm/networks.py:1076-1081) writes its 2 / 3 channels planar for the warp kernels; its
    backward runs on the tcgen05 kernels through a channels-last copy of the gradient padded to 4 channels."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    nd = len(shape)
    conv = F.conv2d if nd == 2 else F.conv3d
    r = gi.rng(990 + Cin + Cout + nd)
    x = torch.from_numpy(r.standard_normal((2, Cin, *shape)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, *([3] * nd))) / np.sqrt(Cin * 3 ** nd)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    y = conv(x, w, b, padding=1)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)
    xq, wq, gq = (tp.tf32_round(t.detach()).double() for t in (x, w, gy))
    grad_in = torch.nn.grad.conv2d_input if nd == 2 else torch.nn.grad.conv3d_input
    grad_w = torch.nn.grad.conv2d_weight if nd == 2 else torch.nn.grad.conv3d_weight
    emu_dx, emu_dw = grad_in(x.shape, wq, gq, padding=1), grad_w(xq, w.shape, gq, padding=1)
    prev = Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto"
    prof = Fn.ConvProfile(); Fn.PROFILE = prof
    try:
        xg = x.detach().cuda().movedim(1, -1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        yg = Fn.conv_cl(xg, wg, bg, pad=1, planar_out=True)
        assert yg.shape == y.shape
        yg.backward(gy.cuda())
        torch.cuda.synchronize()
    finally:
        Fn.CONV_ENGINE, Fn.PROFILE = prev, None
    kinds = prof.by_kind()
    assert "umma_dgrad" in kinds and "umma_wgrad" in kinds, kinds.keys()
    for name, got, want, em in (("dgrad", xg.grad.movedim(-1, 1).cpu(), x.grad, emu_dx), ("wgrad", wg.grad.cpu(), w.grad, emu_dw)):
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 3e-3 * scale, name
        assert float((got.double() - em).abs().max()) <= 3e-5 * scale, (name, "vs TF32-truncated float64 reference")
    np.testing.assert_allclose(bg.grad.cpu().numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


def test_patchnce_tensor_core_3xtf32(orc):
    """K6: PatchNCE logits S = Q K^T and dQ = dS K on the tcgen05 kernel (batched, per-sample weight tiles) with
    3xTF32-split operands: must agree with the fp32 CUDA-core path and the C oracle to fp32 accuracy (the reference's
    torch.bmm is fp32), far tighter than plain TF32 would (1e-3 * logits / T)."""
    import dfmir_b200.functional as Fn
    B, P, D = 3, 256, 256
    r = gi.rng(77)
    q = r.standard_normal((B * P, D)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    k = r.standard_normal((B * P, D)).astype(np.float32); k /= np.linalg.norm(k, axis=1, keepdims=True)
    k = (0.6 * q + 0.4 * k).astype(np.float32)            # correlated positives
    gw = r.standard_normal(B * P).astype(np.float32)
    res = {}
    for eng in ("simt", "auto"):
        Fn.CONV_ENGINE = eng
        try:
            qg = torch.from_numpy(q).cuda().requires_grad_()
            n0 = None
            loss = Fn.patchnce(qg, torch.from_numpy(k).cuda(), B, 0.07)
            (loss * torch.from_numpy(gw).cuda()).sum().backward()
            res[eng] = (loss.detach().cpu().numpy(), qg.grad.cpu().numpy())
        finally:
            Fn.CONV_ENGINE = "auto"
    want = orc.patchnce(q, k, B, 0.07)
    np.testing.assert_allclose(res["simt"][0], want, atol=1e-4)
    np.testing.assert_allclose(res["auto"][0], want, atol=1e-4)
    np.testing.assert_allclose(res["auto"][0], res["simt"][0], atol=2e-5)
    np.testing.assert_allclose(res["auto"][1], res["simt"][1], atol=1e-4 * np.abs(res["simt"][1]).max())
    # the split products really ran on the tensor cores: a plain-TF32 product would be ~1e-2 off at T = 0.07
    assert np.abs(res["auto"][0] - res["simt"][0]).max() < 1e-4


@pytest.mark.parametrize("N,Cin,Cout,H,W,pad", [(2, 256, 256, 66, 66, 0), (1, 256, 64, 20, 28, 1), (3, 512, 32, 18, 18, 0)])
def test_dgrad_accumulates_into_prefilled_dx(N, Cin, Cout, H, W, pad):
    """dfmir_conv_umma_dgrad_acc (CTA-pair kernel, accumulating epilogue): dx += conv_transpose(dy, w) on a dx that
    already holds the skip connection's gradient (ResnetBlock, models/networks.py:1218-1221).  Against the float64
    data gradient of TF32-truncated operands plus the pre-filled values, 3e-5 of the scale."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    from dfmir_b200 import _lib
    r = gi.rng(1200 + Cin + Cout + H)
    OH, OW = H + 2 * pad - 2, W + 2 * pad - 2
    dy = torch.from_numpy(r.standard_normal((N, Cout, OH, OW)).astype(np.float32))
    w = torch.from_numpy((r.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(Cout * 9)).astype(np.float32))
    pre = torch.from_numpy(r.standard_normal((N, Cin, H, W)).astype(np.float32))
    want = pre.double() + torch.nn.grad.conv2d_input((N, Cin, H, W), tp.tf32_round(w).double(), tp.tf32_round(dy).double(), padding=pad)
    dyg = dy.cuda().permute(0, 2, 3, 1).contiguous()
    dx = pre.cuda().permute(0, 2, 3, 1).contiguous()
    wk = w.cuda().reshape(Cout, Cin, 9).permute(2, 1, 0).contiguous()          # [tap][Cin][Cout]: K-major for the data gradient
    d = Fn._make_desc(2, N, Cin, Cout, [H, W], [OH, OW], [3, 3], [pad, pad], 1, 0, Fn._cl_strides(dx, 2), Fn._cl_strides(dyg, 2))
    assert _lib.lib().dfmir_conv_umma_dgrad_acc_supported(ctypes.byref(d))
    _lib.call("dfmir_conv_umma_dgrad_acc", dyg, wk, dx, ctypes.byref(d))
    got = dx.permute(0, 3, 1, 2).cpu().double()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 3e-5 * scale
    # the epilogue added to dx rather than overwriting it
    assert float((got - pre.double()).abs().max()) > 1e-2


def test_resnet_block_slot_handovers():
    """The mutable side channels between the autograd nodes of a ResnetBlock (functional.BiasGradSlot,
    ResidualGradSlot; networks.ResnetBlock.forward_padded): conv bias gradients summed by the InstanceNorm backward,
    skip-connection gradient parked by the last norm's backward and completed by the first convolution's
    accumulating data-gradient epilogue.  Two chained blocks at 256 channels against the float64 autograd of the
    reference's module graph with TF32-truncated convolution operands (oracle/torch_port.py): every gradient at
    accumulation-noise level, and the slot path equal to the slot-free path."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    from dfmir_b200 import networks
    import torch.nn as nn
    import functools
    r = gi.rng(1300)
    N, C, Hh = 2, 256, 24
    norm = functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    blocks = [networks.ResnetBlock(C, 'reflect', norm, False, True) for _ in range(2)]
    for b in blocks:
        for conv in (b.conv_block[1], b.conv_block[5]):
            conv.weight.data = torch.from_numpy((r.standard_normal((C, C, 3, 3)) / np.sqrt(C * 9)).astype(np.float32))
            conv.bias.data = torch.from_numpy(r.standard_normal(C).astype(np.float32) * 0.1)
    x = torch.from_numpy(r.standard_normal((N, C, Hh, Hh)).astype(np.float32))
    gy = torch.from_numpy(r.standard_normal((N, C, Hh, Hh)).astype(np.float32))

    # float64 reference with truncated operands
    tp.TF32_EMULATION = "trunc"
    try:
        xr = x.double().requires_grad_()
        leaves = []
        a = xr
        for b in blocks:
            w1, b1 = b.conv_block[1].weight.detach().double().requires_grad_(), b.conv_block[1].bias.detach().double().requires_grad_()
            w2, b2 = b.conv_block[5].weight.detach().double().requires_grad_(), b.conv_block[5].bias.detach().double().requires_grad_()
            leaves += [w1, b1, w2, b2]
            h = F.relu(F.instance_norm(tp.conv2d(F.pad(a, (1,) * 4, mode='reflect'), w1, b1)))
            h = F.instance_norm(tp.conv2d(F.pad(h, (1,) * 4, mode='reflect'), w2, b2))
            a = a + h
        (a * gy.double()).sum().backward()
    finally:
        tp.TF32_EMULATION = None

    def run(use_slots):
        for b in blocks:
            b.cuda()
            b.zero_grad()
        xg = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        P = Fn.pad_reflect_cl(xg, 1)
        for i, b in enumerate(blocks):
            op = 1 if i == 0 else 0
            if use_slots:
                P = b.forward_padded(P, op)
            else:
                c1, c2 = b.conv_block[1], b.conv_block[5]
                y = Fn.conv_cl(P, c1.weight, c1.bias)
                P1 = Fn.instnorm_cl(y, relu=True, out_pad=1)
                y = Fn.conv_cl(P1, c2.weight, c2.bias)
                P = Fn.instnorm_cl(y, relu=False, out_pad=op, res=P, res_pad=1)
        (P * gy.cuda().permute(0, 2, 3, 1)).sum().backward()
        params = [p.grad.detach().cpu().double() for b in blocks for p in (b.conv_block[1].weight, b.conv_block[1].bias,
                                                                         b.conv_block[5].weight, b.conv_block[5].bias)]
        return P.detach().permute(0, 3, 1, 2).cpu().double(), xg.grad.permute(0, 3, 1, 2).cpu().double(), params

    prev_engine, prev_min = Fn.CONV_ENGINE, Fn.UMMA_MIN_POSITIONS
    Fn.CONV_ENGINE, Fn.UMMA_MIN_POSITIONS = "auto", 0
    try:
        out_s, dx_s, gr_s = run(True)
        out_n, dx_n, gr_n = run(False)
    finally:
        Fn.CONV_ENGINE, Fn.UMMA_MIN_POSITIONS = prev_engine, prev_min
    # chained TF32 layers: an activation within rounding of a truncation boundary may truncate differently in the two
    # pipelines (a 2^-11 relative step in one operand), so pipelines agree to ~1e-4, not to accumulation noise
    assert float((out_s - a.detach()).abs().max()) <= 5e-4 * float(a.detach().abs().max())
    wscale = max(float(l.grad.abs().max()) for l in leaves[0::2])
    for name, got_s, got_n, want in [("dx", dx_s, dx_n, xr.grad)] + [(f"param{i}", gs, gn, l.grad) for i, (gs, gn, l) in enumerate(zip(gr_s, gr_n, leaves))]:
        sc = float(want.abs().max())
        if name.startswith("param") and int(name[5:]) % 2 == 1:
            sc = wscale       # conv biases in front of an instance norm: the true gradient is zero, compare on the weights' scale
        # vs the reference pipeline: TF32 truncation boundaries make two pipelines agree in norm, not element by element
        if not (name.startswith("param") and int(name[5:]) % 2 == 1):
            assert float((got_s - want).norm() / want.norm()) <= 2e-2, (name, "slots vs truncated float64", float((got_s - want).norm() / want.norm()))
        assert float((got_s - got_n).abs().max()) <= 5e-4 * sc, (name, "slot path vs slot-free path", float((got_s - got_n).abs().max()), sc)


def test_sparse_tap_gradient_equals_dense():
    """PatchNCE taps (functional.gather_patches): the sparse COO gradient of a tapped activation that autograd
    index-adds into the dense gradient arriving from the next layer must equal the dense scatter path."""
    import dfmir_b200.functional as Fn
    r = gi.rng(1400)
    B, H, W, C, P = 3, 20, 24, 32, 64
    x = torch.from_numpy(r.standard_normal((B, H, W, C)).astype(np.float32)).cuda()
    ids = torch.from_numpy(r.permutation(H * W)[:P]).cuda()
    gw = torch.from_numpy(r.standard_normal((B * P, C)).astype(np.float32)).cuda()
    gd = torch.from_numpy(r.standard_normal((B, H, W, C)).astype(np.float32)).cuda()
    res = {}
    for sparse in (True, False):
        prev = Fn.SPARSE_TAP_GRAD
        Fn.SPARSE_TAP_GRAD = sparse
        try:
            xl = x.clone().requires_grad_()
            y = xl * 1.0                               # a non-leaf channels-last activation, as in ResnetGenerator.forward
            v = y.permute(0, 3, 1, 2)
            v._dfmir_cl = y
            rows = Fn.gather_patches(v, ids)
            ((rows * gw).sum() + (y * gd).sum()).backward()      # tap gradient + the next layer's dense gradient
            res[sparse] = (rows.detach().clone(), xl.grad.clone())
        finally:
            Fn.SPARSE_TAP_GRAD = prev
    assert torch.equal(res[True][0], res[False][0])
    want = gd.clone().view(B, H * W, C)
    want[:, ids, :] += gw.view(B, P, C)
    assert torch.equal(res[False][1], want.view(B, H, W, C))
    assert torch.allclose(res[True][1], res[False][1], rtol=0, atol=1e-6)


def test_sparse_tap_gradient_on_padded_buffer_equals_dense():
    """The same for a tap that is the interior of a reflect-padded buffer (the ResnetBlock outputs, layers 12 / 16 of the
    generator): sparse gradient on the padded buffer itself vs the dense path through the slice's backward."""
    import dfmir_b200.functional as Fn
    r = gi.rng(1401)
    B, H, W, C, P, pad = 2, 12, 20, 16, 48, 1
    x = torch.from_numpy(r.standard_normal((B, H + 2 * pad, W + 2 * pad, C)).astype(np.float32)).cuda()
    ids = torch.from_numpy(r.permutation(H * W)[:P]).cuda()
    gw = torch.from_numpy(r.standard_normal((B * P, C)).astype(np.float32)).cuda()
    gd = torch.from_numpy(r.standard_normal(tuple(x.shape)).astype(np.float32)).cuda()
    res = {}
    for sparse in (True, False):
        prev = Fn.SPARSE_TAP_GRAD
        Fn.SPARSE_TAP_GRAD = sparse
        try:
            xl = x.clone().requires_grad_()
            Pb = xl * 1.0
            v = Pb[:, pad:H + pad, pad:W + pad, :].permute(0, 3, 1, 2)
            v._dfmir_cl_pad = (Pb, pad)
            rows = Fn.gather_patches(v, ids)
            ((rows * gw).sum() + (Pb * gd).sum()).backward()
            res[sparse] = (rows.detach().clone(), xl.grad.clone())
        finally:
            Fn.SPARSE_TAP_GRAD = prev
    assert torch.equal(res[True][0], res[False][0])
    want = gd.clone()
    inner = want[:, pad:H + pad, pad:W + pad, :].reshape(B, H * W, C).clone()
    inner[:, ids, :] += gw.view(B, P, C)
    want[:, pad:H + pad, pad:W + pad, :] = inner.view(B, H, W, C)
    assert torch.allclose(res[False][1], want, rtol=0, atol=1e-6)
    assert torch.allclose(res[True][1], want, rtol=0, atol=1e-6)


@pytest.mark.parametrize("nd,N,Cin,Cout,S", [(3, 1, 2, 16, (16, 24, 32)), (3, 2, 16, 32, (8, 16, 24)), (3, 1, 32, 64, (8, 8, 16)),
                                             (2, 2, 2, 16, (64, 48)), (2, 1, 16, 32, (32, 40)), (2, 2, 64, 64, (16, 24))])
def test_strided_encoder_conv_on_tensor_cores(nd, N, Cin, Cout, S):
    """VoxelMorph's stride-2 encoder convolutions (vxm networks.py:1514-1515 with stride=2) on the tcgen05 engine: the
    layer runs as a stride-1 2^nd convolution over the space-to-depth activation (functional.conv_cl, S2D_STRIDED).
    Forward, data gradient and weight gradient against the float64 stride-2 convolution of TF32-truncated operands (3e-5
    of the scale) and the exact fp32 result (3e-3)."""
    from oracle import torch_port as tp
    import dfmir_b200.functional as Fn
    conv = F.conv2d if nd == 2 else F.conv3d
    grad_in = torch.nn.grad.conv2d_input if nd == 2 else torch.nn.grad.conv3d_input
    grad_w = torch.nn.grad.conv2d_weight if nd == 2 else torch.nn.grad.conv3d_weight
    r = gi.rng(1500 + Cin + Cout + nd)
    x = torch.from_numpy(r.standard_normal((N, Cin, *S)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, *([3] * nd))) / np.sqrt(Cin * 3 ** nd)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    with torch.no_grad():
        y = F.leaky_relu(conv(x, w, b, stride=2, padding=1), 0.2)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    xq, wq = tp.tf32_round(x.detach()).double(), tp.tf32_round(w.detach()).double()
    emu_y = F.leaky_relu(conv(xq, wq, b.detach().double(), stride=2, padding=1), 0.2)
    prev = Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto"
    prof = Fn.ConvProfile(); Fn.PROFILE = prof
    try:
        xg = x.detach().cuda().movedim(1, -1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        yg = Fn.conv_cl(xg, wg, bg, stride=2, pad=1, act=Fn.ACT_LEAKY)
        assert tuple(yg.shape) == (N, *[s // 2 for s in S], Cout)
        yg.backward(gy.cuda().movedim(1, -1).contiguous())
        torch.cuda.synchronize()
    finally:
        Fn.CONV_ENGINE, Fn.PROFILE = prev, None
    # the activation's backward uses the sign of the kernel's own output (an output within rounding of zero may differ
    # in sign from the reference's): references of the two backward products take the same mask
    gpre = gy * torch.where(yg.detach().movedim(-1, 1).cpu() > 0, 1.0, 0.2)
    gq = tp.tf32_round(gpre).double()
    emu_dx = grad_in(x.shape, wq, gq, stride=2, padding=1)
    emu_dw = grad_w(xq, w.shape, gq, stride=2, padding=1)
    x.grad = grad_in(x.shape, w.detach(), gpre, stride=2, padding=1)
    w.grad = grad_w(x.detach(), w.shape, gpre, stride=2, padding=1)
    b.grad = gpre.sum(dim=[0] + list(range(2, nd + 2)))
    kinds = prof.by_kind()
    # the weight gradient needs >= 16 channels on both sides (2-D first layer: 4 x 2 = 8 -> exact fp32 kernel)
    wgrad_tc = (1 << nd) * Cin >= 16
    assert set(kinds) == {"umma_fwd", "umma_dgrad", "umma_wgrad" if wgrad_tc else "simt"}, kinds.keys()
    for name, got, want, em in (("fwd", yg.detach().movedim(-1, 1).cpu(), y.detach(), emu_y),
                                ("dgrad", xg.grad.movedim(-1, 1).cpu(), x.grad, emu_dx), ("wgrad", wg.grad.cpu(), w.grad, emu_dw)):
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 3e-3 * scale, name
        if name == "wgrad" and not wgrad_tc:
            assert float((got - want).abs().max()) <= 2e-4 * scale
            continue
        assert float((got.double() - em).abs().max()) <= 3e-5 * scale, (name, "vs TF32-truncated float64 reference")
    np.testing.assert_allclose(bg.grad.cpu().numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


@pytest.mark.parametrize("N,Cin,Cout,H,W,pad", [(2, 256, 256, 34, 34, 0), (3, 64, 128, 40, 56, 1), (2, 128, 64, 24, 72, 1),
                                                (1, 128, 256, 21, 13, 1), (2, 256, 128, 16, 16, 1)])
def test_instnorm_statistics_from_conv_epilogue(N, Cin, Cout, H, W, pad):
    """InstanceNorm statistics as a by-product of the tcgen05 forward epilogue (dfmir_conv_umma_fwd_stats ->
    dfmir_instnorm_fwd_rows; CTA-pair and halo kernels, full and partial tiles): the normalised output and its
    backward equal the path that reads the convolution output again (dfmir_instnorm_fwd) to summation-order noise."""
    import dfmir_b200.functional as Fn
    r = gi.rng(1600 + Cin + Cout + H)
    x = torch.from_numpy(r.standard_normal((N, H, W, Cin)).astype(np.float32)).cuda()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, 3, 3)) / np.sqrt(Cin * 9)).astype(np.float32)).cuda()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).cuda()
    OH, OW = H + 2 * pad - 2, W + 2 * pad - 2
    gy = torch.from_numpy(r.standard_normal((N, OH + 2, OW + 2, Cout)).astype(np.float32)).cuda()
    prev = Fn.CONV_ENGINE
    Fn.CONV_ENGINE = "auto"
    out = {}
    try:
        for fused in (True, False):
            xg = x.clone().requires_grad_()
            st = Fn.StatsSlot() if fused else None
            y = Fn.conv_cl(xg, w, b, pad=pad, stats_slot=st)
            if fused:
                assert st.rows is not None and st.rows.shape[0] == N and st.rows.shape[2] == Cout, "the epilogue did not produce statistic rows"
                # the row sums add up to the plain per-channel sums of the stored result
                s = st.rows.double().sum(dim=1)
                ref = torch.stack([y.detach().double().sum(dim=(1, 2)), (y.detach().double() ** 2).sum(dim=(1, 2))], dim=-1)
                assert float((s - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
            a = Fn.instnorm_cl(y, relu=True, out_pad=1, stats_slot=st)
            a.backward(gy)
            out[fused] = (a.detach(), xg.grad)
    finally:
        Fn.CONV_ENGINE = prev
    assert float((out[True][0] - out[False][0]).abs().max()) <= 2e-5
    # backward: the norm's output gradient feeds a TF32 data-gradient product; 1e-7 differences in the statistics move
    # a few operands across a truncation boundary (2^-11 relative each)
    assert float((out[True][1] - out[False][1]).abs().max()) <= 2e-4 * float(out[False][1].abs().max()) + 1e-7


@pytest.mark.parametrize("M,K,N,relu", [(4096, 256, 256, True), (4096, 128, 256, True), (8192, 256, 256, False)])
def test_linear_on_tensor_cores_is_fp32_class(M, K, N, relu):
    """PatchSampleF's nn.Linear layers (models/networks.py:587-595) at >= 4096 rows run on the tcgen05 kernels with
    3xTF32-split operands: forward, dx, dW and db agree with float64 to fp32 accuracy (1e-5 of the scale; plain TF32
    would be ~1e-3), because the reference's nn.Linear is an fp32 product."""
    import dfmir_b200.functional as Fn
    r = gi.rng(1700 + K + N)
    x = torch.from_numpy(r.standard_normal((M, K)).astype(np.float32))
    W = torch.from_numpy((r.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
    b = torch.from_numpy(r.standard_normal(N).astype(np.float32))
    gy = torch.from_numpy(r.standard_normal((M, N)).astype(np.float32))
    xr, Wr, br = x.double().requires_grad_(), W.double().requires_grad_(), b.double().requires_grad_()
    yr = torch.nn.functional.linear(xr, Wr, br)
    if relu:
        yr = torch.relu(yr)
    yr.backward(gy.double())
    prof = Fn.ConvProfile(); Fn.PROFILE = prof
    try:
        xg, Wg, bg = x.cuda().requires_grad_(), W.cuda().requires_grad_(), b.cuda().requires_grad_()
        y = Fn.linear(xg, Wg, bg, relu=relu)
        y.backward(gy.cuda())
        torch.cuda.synchronize()
    finally:
        Fn.PROFILE = None
    assert set(prof.by_kind()) == {"umma_fwd", "umma_dgrad", "umma_wgrad"}, prof.by_kind().keys()
    for name, got, want in (("y", y.detach(), yr.detach()), ("dx", xg.grad, xr.grad), ("dW", Wg.grad, Wr.grad), ("db", bg.grad, br.grad)):
        sc = float(want.abs().max())
        err = float((got.cpu().double() - want).abs().max())
        assert err <= 1e-5 * sc, (name, err, sc)
