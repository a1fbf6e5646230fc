"""CUDA-graph capture of the training step (REGISTRATIONModel.capture_step): the replayed step trains
(parameters move, losses stay finite and close to the eager trajectory) and reads the inputs set_input() copies into
the captured buffers."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(cuda_graph):
    import bench
    from dfmir_b200 import registration_model as rm
    opt = rm.default_options(batch_size=2, crop_size=64, load_size=64, gpu_ids=[0], cuda_graph=cuda_graph)
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rm.REGISTRATIONModel(opt)
        A, B = bench.synthetic_pair(2, 64, 7)
        data = {"A": A, "B": B}
        model.data_dependent_initialize(data)
        model.setup(opt)
    model.set_input(data)
    return model, data


def test_captured_step_trains():
    from dfmir_b200 import _lib
    model, data = make(True)
    for _ in range(3):
        model.optimize_parameters()
    eager = model.get_current_losses()
    model.capture_step()
    assert model.graph_launches_per_step > 100
    w0 = [p.detach().clone() for p in model.netG.parameters()][:4] + [p.detach().clone() for p in model.netR.parameters()][:2]
    n0 = _lib.launch_count()
    for _ in range(3):
        model.set_input(data)
        model.optimize_parameters()
    assert _lib.launch_count() == n0, "replays must not launch from the host"
    got = model.get_current_losses()
    assert all(np.isfinite(v) for v in got.values()), got
    for k in ("NCE", "R"):
        assert abs(got[k] - eager[k]) <= 0.3 * abs(eager[k]), (k, got[k], eager[k])
    w1 = [p.detach() for p in model.netG.parameters()][:4] + [p.detach() for p in model.netR.parameters()][:2]
    assert all(float((a - b).abs().max()) > 0 for a, b in zip(w0, w1)), "parameters did not move under replay"
    # different inputs through the captured buffers change the losses
    model.set_input({"A": data["B"], "B": data["A"]})
    model.optimize_parameters()
    swapped = model.get_current_losses()
    assert abs(swapped["R"] - got["R"]) > 1e-6


def test_capture_needs_capturable_optimizers():
    from dfmir_b200 import _lib
    model, _ = make(False)
    with pytest.raises(_lib.DfmirError):
        model.capture_step()
