"""Fused multi-tensor Adam (csrc/adam.cu, dfmir_b200/optim.py) against torch.optim.Adam - the optimizer of the
reference's step (models/registration_model.py:114-115, 135, 168-171: betas (0.5, 0.999), eps 1e-8)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 1, 7, 7), (64,), (256, 256, 3, 3), (3,), (5, 7), (16, 36, 3, 3, 3), (1,), (4099,)]


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).cuda().requires_grad_() for s in SHAPES]


@pytest.mark.parametrize("lr_tensor", [False, True])
def test_fused_adam_matches_torch_adam(lr_tensor):
    from dfmir_b200.optim import FusedAdam
    pa, pb = _make(1), _make(1)
    # an unaligned parameter / gradient (views at a 4-byte offset): the scalar path of the kernel
    base_a, base_b = torch.randn(1001).cuda(), None
    base_b = base_a.clone()
    ua, ub = base_a[1:].detach().requires_grad_(), base_b[1:].detach().requires_grad_()
    pa.append(ua); pb.append(ub)
    lr_a = torch.tensor(2e-4, device="cuda") if lr_tensor else 2e-4
    lr_b = torch.tensor(2e-4, device="cuda") if lr_tensor else 2e-4
    oa = FusedAdam(pa, lr=lr_a, betas=(0.5, 0.999))
    ob = torch.optim.Adam(pb, lr=lr_b, betas=(0.5, 0.999), capturable=lr_tensor)
    g = torch.Generator().manual_seed(7)
    for it in range(12):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it % 3 - 1))
            a.grad = gr.clone(); b.grad = gr.clone()
        if it == 6 and lr_tensor:            # a scheduler step: filled in place, as update_learning_rate does
            lr_a.fill_(1e-4); lr_b.fill_(1e-4)
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7), float((a - b).abs().max())
        sa, sb = oa.state[a], ob.state[b]
        assert torch.allclose(sa['exp_avg'], sb['exp_avg'], rtol=1e-5, atol=2e-6)      # lerp rounding where g and m cancel
        assert torch.allclose(sa['exp_avg_sq'], sb['exp_avg_sq'], rtol=1e-5, atol=1e-9)
        assert float(sa['step']) == float(sb['step']) == 12.0


def test_fused_adam_state_dict_interchanges_with_torch_adam():
    """A torch.optim.Adam checkpoint (base_model.py-style optimizer state) continues under FusedAdam and back."""
    from dfmir_b200.optim import FusedAdam
    pa, pb = _make(3), _make(3)
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(9)
    grads = [[torch.randn(p.shape, generator=g).cuda() for p in pa] for _ in range(6)]
    for it in range(3):
        for b, gr in zip(pb, grads[it]):
            b.grad = gr.clone()
        ob.step()
    for a, b in zip(pa, pb):
        a.data.copy_(b.data)
    oa = FusedAdam(pa, lr=2e-4, betas=(0.5, 0.999))
    oa.load_state_dict(copy.deepcopy(ob.state_dict()))      # (load_state_dict shares same-device tensors with its argument)
    for it in range(3, 6):
        for a, b, gr in zip(pa, pb, grads[it]):
            a.grad = gr.clone(); b.grad = gr.clone()
        oa.step(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)
    oc = torch.optim.Adam(_make(3), lr=2e-4, betas=(0.5, 0.999))
    oc.load_state_dict(copy.deepcopy(oa.state_dict()))       # and back
    assert float(next(iter(oc.state.values()))['step']) == 6.0


def test_fused_adam_in_cuda_graph():
    from dfmir_b200.optim import FusedAdam
    pa, pb = _make(5), _make(5)
    lr = torch.tensor(2e-4, device="cuda")
    oa = FusedAdam(pa, lr=lr, betas=(0.5, 0.999))
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    static = [torch.zeros_like(p) for p in pa]
    for a, s in zip(pa, static):
        a.grad = s
    oa.step()                                   # eager: builds the pointer table
    for b in pb:
        b.grad = torch.zeros_like(b)
    ob.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        oa.step()
    g = torch.Generator().manual_seed(11)
    for it in range(4):
        for s, b in zip(static, pb):
            gr = torch.randn(s.shape, generator=g).cuda()
            s.copy_(gr); b.grad = gr.clone()
        graph.replay(); ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)


def test_fused_adam_step_invalidates_cached_weight_layouts():
    """The kernel writes the parameters behind autograd's back: step() bumps their version counters, so a kernel-layout
    copy cached by a no-grad forward (functional.packed_weight, keyed on the version) is rebuilt after an update."""
    from dfmir_b200 import functional as Fn
    from dfmir_b200.optim import FusedAdam
    w = torch.randn(8, 4, 3, 3, device="cuda").requires_grad_()
    frozen = torch.randn(5, device="cuda").requires_grad_()        # no gradient: untouched, version unchanged
    opt = FusedAdam([w, frozen], lr=1e-2, betas=(0.5, 0.999))
    with torch.no_grad():
        before = Fn.packed_weight(w)
        assert Fn.packed_weight(w) is before
    v0, f0 = w._version, frozen._version
    w.grad = torch.ones_like(w)
    opt.step()
    assert w._version > v0 and frozen._version == f0
    with torch.no_grad():
        after = Fn.packed_weight(w)
    assert after is not before
    assert torch.equal(after, w.detach().reshape(8, 4, 9).permute(2, 1, 0).contiguous())
    assert not torch.equal(after, before)
