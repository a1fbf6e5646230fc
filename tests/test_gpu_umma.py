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
    np.testing.assert_allclose(tc[2].numpy(), w.grad.numpy(), atol=2e-4 * float(w.grad.abs().max()))
    np.testing.assert_allclose(tc[3].numpy(), b.grad.numpy(), atol=2e-4 * float(b.grad.abs().max()))


def test_umma_supported_shapes():
    import dfmir_b200.umma as umma
    x = torch.zeros(1, 8, 8, 64, device="cuda")
    assert umma.supported(2, 64, 128, [3, 3], 1, [1, 1], x, False)
    assert not umma.supported(2, 1, 64, [7, 7], 1, [0, 0], x, False)       # stem
    assert not umma.supported(2, 64, 1, [7, 7], 1, [0, 0], x, False)       # head
    assert not umma.supported(2, 64, 128, [3, 3], 2, [1, 1], x, False)     # strided
    assert not umma.supported(3, 64, 128, [3, 3, 3], 1, [1, 1, 1], x, False)
