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

CASES = [  # N, Cin, Cout, H, W, k, pad   (H, W = input spatial size)
    (2, 64, 128, 16, 128, 3, 1),     # TW = 128, TH = 1
    (2, 128, 256, 8, 64, 3, 1),      # TW = 64, TH = 2
    (1, 256, 256, 18, 18, 3, 0),     # ResnetBlock conv on a reflect-padded buffer: 16x16 output, TW = 16
    (2, 256, 128, 12, 40, 3, 1),     # partial tiles in w
    (1, 128, 64, 9, 256, 3, 1),      # two tiles per row
    (3, 64, 64, 10, 10, 1, 0),       # 1x1
    (1, 256, 256, 66, 66, 3, 0),     # full-size ResnetBlock geometry: dgrad output 66x66 (partial tiles)
]


@pytest.mark.parametrize("case", CASES, ids=[f"u{i}" for i in range(len(CASES))])
def test_umma_conv_fwd_dgrad(case):
    from dfmir_b200 import _lib
    import dfmir_b200.functional as Fn
    N, Cin, Cout, H, W, k, pad = case
    r = gi.rng(900 + Cin + Cout + H)
    x = torch.from_numpy(r.standard_normal((N, Cin, H, W)).astype(np.float32)).requires_grad_()
    w = torch.from_numpy((r.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)).requires_grad_()
    b = torch.from_numpy(r.standard_normal(Cout).astype(np.float32)).requires_grad_()
    y = F.conv2d(x, w, b, padding=pad)
    gy = torch.from_numpy(r.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(gy)

    def run(engine):
        Fn.CONV_ENGINE = engine
        xg = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_()
        wg, bg = w.detach().cuda().requires_grad_(), b.detach().cuda().requires_grad_()
        n0 = _lib.launch_count()
        yg = Fn.conv_cl(xg, wg, bg, pad=pad)
        yg.backward(gy.cuda().permute(0, 2, 3, 1).contiguous())
        torch.cuda.synchronize()
        return yg.detach().permute(0, 3, 1, 2).cpu(), xg.grad.permute(0, 3, 1, 2).cpu(), wg.grad.cpu(), bg.grad.cpu()

    try:
        exact = run("simt")
        tc = run("auto")
    finally:
        Fn.CONV_ENGINE = "auto"
    for name, got, want, ex in (("fwd", tc[0], y.detach(), exact[0]), ("dgrad", tc[1], x.grad, exact[1])):
        scale = float(want.abs().max())
        assert float((ex - want).abs().max()) <= 3e-5 * scale, name + " (fp32 engine)"
        err = float((got - want).abs().max())
        assert err <= 3e-3 * scale, (name, err, scale)
        assert err > 0 or name == "dgrad", "tensor-core result is bit-identical to fp32: engine did not run?"
    # weight gradient: tcgen05 split-K kernel where supported (TF32 operands), else the fp32 kernel
    np.testing.assert_allclose(exact[2].numpy(), w.grad.numpy(), atol=2e-4 * float(w.grad.abs().max()))
    np.testing.assert_allclose(tc[2].numpy(), w.grad.numpy(), atol=3e-3 * float(w.grad.abs().max()))
    np.testing.assert_allclose(tc[3].numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


def test_umma_supported_shapes():
    import dfmir_b200.umma as umma
    x = torch.zeros(1, 8, 8, 64, device="cuda")
    assert umma.supported(2, 64, 128, [3, 3], 1, [1, 1], x, False)
    assert not umma.supported(2, 1, 64, [7, 7], 1, [0, 0], x, False)       # stem
    assert not umma.supported(2, 64, 1, [7, 7], 1, [0, 0], x, False)       # head
    assert not umma.supported(2, 64, 128, [3, 3], 2, [1, 1], x, False)     # strided
    assert not umma.supported(3, 64, 128, [3, 3, 3], 1, [1, 1, 1], x, False)


def test_generator_ngf64_tensor_core_vs_cpu_port():
    """ResnetGenerator at the reference width (ngf = 64, all five conv shapes on the tcgen05 engine, fwd +
    dgrad + wgrad) against the torch CPU fp32 port of the reference module (oracle/torch_port.py).
    TF32 operands through 4 residual blocks: output within 2e-2 absolute of a tanh output (|x| <= 1),
    weight gradients within 5e-2 of each tensor's largest entry."""
    from oracle import torch_port as tp
    from dfmir_b200 import networks
    import dfmir_b200.functional as Fn
    sdG, _, _ = tp.random_state_dicts(ngf=64, n_blocks=4, crop=64, seed=3)
    sdG = {k: (v * 8.0 if k.endswith("weight") else v) for k, v in sdG.items()}   # gain 0.16: a non-trivial output
    x = torch.from_numpy(gi.image_textured(411, 2, (64, 64)))
    wts = torch.from_numpy(gi.weights(412, (2, 1, 64, 64), 1.0))
    leaves = {k: v.clone().requires_grad_() for k, v in sdG.items()}
    ref = tp.resnet_generator(x, leaves, 4)
    (ref * wts).sum().backward()

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
    err = float((out.detach().cpu() - ref.detach()).abs().max())
    assert err <= 2e-2, err
    last = max(int(k.split('.')[1]) for k in sdG)
    for k, p in G.named_parameters():
        want = leaves[k].grad
        if k.endswith("bias") and not k.startswith(f"model.{last}."):   # feeds an InstanceNorm: true gradient is zero
            scale = float(leaves[k[:-4] + "weight"].grad.abs().max())
        else:
            scale = float(want.abs().max())
        e = float((p.grad.cpu() - want).abs().max())
        assert e <= 5e-2 * max(scale, 1e-8), (k, e, scale)
