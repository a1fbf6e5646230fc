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


def test_replays_draw_fresh_patch_ids_and_follow_the_lr_schedule():
    """Every replay of the captured step draws new PatchSampleF positions (the Philox offset of the generator is a
    graph input), and update_learning_rate() keeps the graph: capturable Adam reads the rate from a device tensor that
    the scheduler fills in place (the reference's train.py steps the schedulers once per epoch)."""
    model, data = make(True)
    for _ in range(2):
        model.optimize_parameters()
    model.capture_step()
    model.optimize_parameters()
    ids1 = [t.clone() for t in model._last_patch_ids]
    model.optimize_parameters()
    ids2 = [t.clone() for t in model._last_patch_ids]
    assert all(a.shape == b.shape for a, b in zip(ids1, ids2))
    assert any(not torch.equal(a, b) for a, b in zip(ids1, ids2)), "two replays drew the same patch ids"
    lr = model.optimizers[0].param_groups[0]['lr']
    assert torch.is_tensor(lr) and lr.is_cuda
    model.opt.n_epochs, model.opt.n_epochs_decay = 0, 3          # linear decay from the first epoch on
    model.schedulers = [__import__('dfmir_b200').networks.get_scheduler(o, model.opt) for o in model.optimizers]
    graph = model._graph
    with contextlib.redirect_stdout(io.StringIO()):
        model.update_learning_rate()
    assert model._graph is graph, "the captured step must survive a learning-rate update"
    new_lr = float(model.optimizers[0].param_groups[0]['lr'])
    assert 0 < new_lr < 2e-4
    # Adam's first-moment direction is unchanged by the rate, so the parameter update scales with it
    p = next(model.netR.parameters())
    before = p.detach().clone()
    model.optimize_parameters()
    step_small = float((p.detach() - before).abs().max())
    assert 0 < step_small <= new_lr * 1.5, (step_small, new_lr)


def test_no_grad_forward_after_replay_uses_current_weights():
    """Kernel-layout weight copies are cached by parameter version; graph replays update the parameters without bumping
    it, so optimize_parameters() drops the caches after every replay: a no-grad forward (model.test())
    between replays must see the current weights."""
    from dfmir_b200 import functional as Fn, umma
    model, data = make(True)
    for _ in range(2):
        model.optimize_parameters()
    model.capture_step()
    model.optimize_parameters()
    model.test()                                   # caches kernel-layout copies under no_grad
    for _ in range(3):
        model.optimize_parameters()
    model.test()
    got = model.fake_B.clone()
    Fn._pack_cache.clear(); umma._kmajor_cache.clear()
    model.test()
    assert torch.equal(got, model.fake_B), "a no-grad forward after graph replays ran on stale weight copies"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2); recorded run: profiles/r2_ddp_equiv.txt")
def test_two_ranks_equal_one_rank_with_the_whole_batch():
    """2 ranks x batch 2 == 1 rank x batch 4 (losses, averaged gradients, parameters after Adam): tools/check_ddp_equiv.py
    under torchrun (reference: nn.DataParallel splits one batch, models/base_model.py:103-107)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "check_ddp_equiv.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "EQUAL" in r.stdout
