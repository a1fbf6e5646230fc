"""Host logic of functional.packed_weight (kernel-layout weight copies shared by the passes of a step): pure torch,
runs on the CPU."""
import torch


def test_packed_weight_layout_grad_and_cache():
    import dfmir_b200.functional as Fn
    Fn._pack_cache.clear()
    torch.manual_seed(0)
    w = torch.randn(5, 3, 3, 3, requires_grad=True)
    p1 = Fn.packed_weight(w)
    assert p1.shape == (9, 3, 5)
    assert torch.equal(p1.detach(), w.detach().reshape(5, 3, 9).permute(2, 1, 0))
    assert Fn.packed_weight(w) is p1, "second use inside one tape must share the copy"
    pz = Fn.packed_weight(w, 1)                      # zero rows for a zero-padded activation channel
    assert pz is not p1 and pz.shape == (9, 4, 5) and float(pz.detach()[:, 3].abs().max()) == 0.0
    Fn._pack_cache.clear()
    p1 = Fn.packed_weight(w)
    g = torch.randn(9, 3, 5)
    (p1 * g).sum().backward()                        # two uses would accumulate into the same node
    assert torch.allclose(w.grad, g.permute(2, 1, 0).reshape(5, 3, 3, 3))
    assert id(w) not in Fn._pack_cache, "the backward pass retires the copy"
    p2 = Fn.packed_weight(w)
    assert p2 is not p1
    with torch.no_grad():
        w.add_(1.0)                                  # what an optimizer step does: bumps the version counter
    p3 = Fn.packed_weight(w)
    assert p3 is not p2 and torch.equal(p3.detach(), w.detach().reshape(5, 3, 9).permute(2, 1, 0))
    with torch.no_grad():
        p4 = Fn.packed_weight(w)                     # a no-grad pass must not hand out a copy that carries a tape
        assert not p4.requires_grad
    assert Fn.packed_weight(w).requires_grad


def test_packed_weight_purges_dead_entries():
    import dfmir_b200.functional as Fn
    Fn._pack_cache.clear()
    for _ in range(300):
        Fn.packed_weight(torch.randn(2, 2, 1, 1))    # temporaries (linear() passes views): their entries die with them
    assert len(Fn._pack_cache) <= 260


def test_fused_adam_rejects_cpu_tensors():
    """dfmir_b200.optim.FusedAdam has no CPU path: parameters on the host raise DfmirError at step() (no silent fallback)."""
    import pytest
    import torch
    from dfmir_b200 import _lib
    from dfmir_b200.optim import FusedAdam
    p = torch.zeros(8, requires_grad=True)
    p.grad = torch.ones(8)
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises((_lib.DfmirError, RuntimeError)):
        opt.step()
