"""Inference path of the reference (test.py:34-90): eval-mode generator + registration forward
(`registration=True`) and the nearest-neighbour label warp, which test.py runs on the CPU (`y_pred2[1].cpu()`):
the warped label map must equal ATen's CPU grid_sample bit for bit (index parity), the registered image ATen's
bilinear result to interpolation rounding."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def aten_cpu_warp(src, flow, mode):
    """SpatialTransformer.forward (models/voxelmorph/torchvoxelmorph/layers.py:30-48) on CPU tensors."""
    shape = flow.shape[2:]
    grid = torch.stack(torch.meshgrid(*[torch.arange(s) for s in shape], indexing="ij")).float()[None]
    new_locs = grid + flow
    for i in range(len(shape)):
        new_locs[:, i, ...] = 2 * (new_locs[:, i, ...] / (shape[i] - 1) - 0.5)
    new_locs = new_locs.permute(0, 2, 3, 1)[..., [1, 0]]
    return F.grid_sample(src, new_locs, align_corners=True, mode=mode)


def test_inference_path_and_label_warp():
    import bench
    from dfmir_b200 import layers, registration_model as rm
    S, B = 128, 2
    opt = rm.default_options(batch_size=B, crop_size=S, load_size=S, gpu_ids=[0])
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rm.REGISTRATIONModel(opt)
        A, Bm = bench.synthetic_pair(B, S, 11)
        data = {"A": A, "B": Bm}
        model.data_dependent_initialize(data)
        model.setup(opt)
    with torch.no_grad():
        model.netR.flow.weight.mul_(3e4)           # N(0, 1e-5) initial flow head: make the deformation a few pixels
    model.eval()
    model.set_input(data)
    model.test()                                    # test.py:47
    assert model.fake_B.shape == (B, 1, S, S) and torch.isfinite(model.fake_B).all()
    with torch.no_grad():
        idt_B = model.netG(model.real_B)            # test.py:77
        y_src, flow = model.netR(model.real_A, model.real_B, registration=True)      # test.py:78
    assert torch.isfinite(idt_B).all() and float(idt_B.abs().max()) <= 1.0
    assert float(flow.abs().max()) > 0.5, "the test needs a deformation of at least half a pixel"
    r = np.random.RandomState(3)
    label = torch.from_numpy(r.randint(0, 6, size=(B, 1, S, S)).astype(np.float32))
    warped = layers.SpatialTransformer([S, S], mode='nearest').cuda()(label.cuda(), flow)          # test.py:80-81
    ref = aten_cpu_warp(label, flow.cpu(), "nearest")
    assert torch.equal(warped.cpu(), ref), "nearest-mode label warp differs from ATen CPU"
    ref_img = aten_cpu_warp(model.real_A.cpu(), flow.cpu(), "bilinear")
    assert float((y_src.cpu() - ref_img).abs().max()) <= 4e-6
