"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from /root/reference) on
the seeded inputs of tests/golden/inputs.py.  Run in the build container only (the GPU box has no
/root/reference); the outputs are committed.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden.py [section ...]      sections: ops nets step   (default: all)
"""
import inspect
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DFMIR_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import inputs as gi  # noqa: E402


def import_reference():
    """The three shims of SURVEY.md 8c; no reference source is modified."""
    inspect.getargspec = lambda f: inspect.getfullargspec(f)[:4]   # modelio.py:14 (removed in py3.11)
    torch.nn.Module.cuda = lambda self, *a, **k: self              # registration_model.py:98,100
    torch.Tensor.cuda = lambda self, *a, **k: self                 # registration_model.py:148
    if REF not in sys.path:
        sys.path.insert(0, REF)


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    meta = {"torch_version": np.array(torch.__version__)}
    np.savez_compressed(path, **arrays, **meta)
    print(f"  {name}.npz  {os.path.getsize(path) / 1024:.0f} KiB")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def gen_ops():
    import_reference()
    import torch.nn.functional as nnf
    from models.voxelmorph.torchvoxelmorph import layers as rl
    from util.losses import NCC_Loss, Grad_Loss
    from models.registration_model import smooothing_loss, REGISTRATIONModel

    # ---- warp: nearest-mode indices (observed through index-ramp sources), linear values, and the
    # normalised grid the reference hands to F.grid_sample (captured, not recomputed)
    captured = {}
    real_gs = nnf.grid_sample

    def spy(src, grid, **kw):
        captured["grid"] = grid.detach().clone()
        return real_gs(src, grid, **kw)

    out = {}
    for name, shape, sigma, seed in gi.WARP_CASES:
        flow = gi.flow(seed, 1, shape, sigma)
        ramps = gi.index_ramps(1, shape)
        img = gi.image(seed + 100, 1, shape)
        st_lin = rl.SpatialTransformer(shape)
        st_nn = rl.SpatialTransformer(shape, mode='nearest')
        rl.nnf.grid_sample = spy
        y_lin = st_lin(t(img), t(flow)).numpy()
        rl.nnf.grid_sample = real_gs
        ngrid = captured["grid"].numpy()          # (1,*S,nd) xyz order
        nd = len(shape)
        ngrid = np.moveaxis(ngrid, -1, 1)[:, ::-1]  # back to (1,nd,*S), ij order
        y_nn = st_nn(t(ramps), t(flow)).numpy()   # sampled index per axis, 0 where out of bounds
        ones = st_nn(torch.ones(1, 1, *shape), t(flow)).numpy()  # 1 in bounds, 0 outside
        out[name + "/nearest_idx"] = y_nn.astype(np.int16)
        out[name + "/nearest_inb"] = ones.astype(np.uint8)
        out[name + "/linear"] = y_lin
        if sigma in (0.0, 1e-5) or name in ("w3d_32x40x48_s3", "w2d_37x53_s2"):
            out[name + "/ngrid"] = np.ascontiguousarray(ngrid)
    for name, shape, seed in gi.HALF_CASES:
        flow = gi.half_integer_flow(seed, 1, shape)
        st_nn = rl.SpatialTransformer(shape, mode='nearest')
        out[name + "/nearest_idx"] = st_nn(t(gi.index_ramps(1, shape)), t(flow)).numpy().astype(np.int16)
        out[name + "/nearest_inb"] = st_nn(torch.ones(1, 1, *shape), t(flow)).numpy().astype(np.uint8)
    save("warp", **out)

    # ---- warp backward (autograd through the reference module)
    out = {}
    for name, shape, sigma, seed in [("b2d", (48, 64), 2.0, 41), ("b3d", (12, 16, 20), 1.5, 42)]:
        nd = len(shape)
        src = t(gi.image(seed, 2, shape, C=2)).requires_grad_()
        flow = t(gi.flow(seed + 1, 2, shape, sigma)).requires_grad_()
        gout = t(gi.weights(seed + 2, (2, 2, *shape), 1.0))
        y = rl.SpatialTransformer(shape)(src, flow)
        y.backward(gout)
        out[name + "/out"] = y.detach().numpy()
        out[name + "/d_src"] = src.grad.numpy()
        out[name + "/d_flow"] = flow.grad.numpy()
    save("warp_bwd", **out)

    # ---- VecInt + ResizeTransform (forward and backward)
    out = {}
    for name, shape, sigma, seed in [("v2d", (128, 128), 8.0, 51), ("v3d", (16, 20, 24), 4.0, 52),
                                     ("v2d_tiny", (128, 128), 1e-3, 53)]:
        nd = len(shape)
        vec = t(gi.smooth_field(gi.rng(seed), (2, nd, *shape), sigma)).requires_grad_()
        gout = t(gi.weights(seed + 2, (2, nd, *shape), 1.0))
        y = rl.VecInt(shape, 7)(vec)
        y.backward(gout)
        out[name + "/out"] = y.detach().numpy()
        out[name + "/d_vec"] = vec.grad.numpy()
        yn = rl.VecInt(shape, 7)(-vec.detach())
        out[name + "/out_neg"] = yn.numpy()
    for name, shape, seed in [("r2d", (64, 96), 61), ("r3d", (16, 20, 24), 62), ("r2d_odd", (37, 53), 63)]:
        nd = len(shape)
        x = t(gi.weights(seed, (2, nd, *shape), 1.0)).requires_grad_()
        down = rl.ResizeTransform(2, nd)(x)
        gd = t(gi.weights(seed + 1, tuple(down.shape), 1.0))
        down.backward(gd)
        out[name + "/down"] = down.detach().numpy()
        out[name + "/down_dx"] = x.grad.numpy().copy()
        x.grad = None
        up = rl.ResizeTransform(0.5, nd)(x)
        gu = t(gi.weights(seed + 2, tuple(up.shape), 1.0))
        up.backward(gu)
        out[name + "/up"] = up.detach().numpy()
        out[name + "/up_dx"] = x.grad.numpy().copy()
    save("vecint_resize", **out)

    # ---- losses
    out = {}
    for name, shape, seed in [("n2d", (64, 64), 71), ("n3d", (24, 28, 32), 72), ("n2d_odd", (45, 70), 73)]:
        nd = len(shape)
        I = t(gi.image_textured(seed, 2, shape)).requires_grad_()
        J = t(gi.image_textured(seed + 1, 2, shape)).requires_grad_()
        crit = NCC_Loss('cpu', kernel_var=[9] * nd, kernel_type='mean')
        loss = crit(I, J)
        loss.backward()
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/cc"] = crit.ncc(I, J).detach().numpy()
        out[name + "/dI"] = I.grad.numpy().copy()
        out[name + "/dJ"] = J.grad.numpy().copy()
        I.grad = None; J.grad = None
        mask = t((gi.image(seed + 2, 2, shape) > -0.5).astype(np.float32))
        lm = crit(I, J, mask=mask)
        lm.backward()
        out[name + "/loss_masked"] = lm.detach().numpy()
        out[name + "/dI_masked"] = I.grad.numpy().copy()
        out[name + "/loss_self"] = crit(I.detach(), I.detach()).numpy()
    # survey known answers: NCC of rand/rand 2x1x64^2 under torch seed 1234
    torch.manual_seed(1234)
    a, b = torch.rand(2, 1, 64, 64), torch.rand(2, 1, 64, 64)
    out["survey/ncc_rand_a"] = a.numpy(); out["survey/ncc_rand_b"] = b.numpy()
    out["survey/ncc_rand"] = NCC_Loss('cpu', kernel_var=[9, 9])(a, b).numpy()
    g = torch.randn(1, 3, 32, 32, 32)
    out["survey/grad_in"] = g.numpy()
    out["survey/grad_l2_3d"] = Grad_Loss(dim=3, penalty='l2')(g).numpy()

    for name, shape, seed in [("g2d", (64, 80), 81), ("g3d", (12, 16, 20), 82)]:
        nd = len(shape)
        for pen in ("l1", "l2"):
            x = t(gi.weights(seed, (2, nd, *shape), 1.0)).requires_grad_()
            loss = Grad_Loss(dim=nd, penalty=pen, loss_mult=2 if pen == "l1" else None)(x)
            loss.backward()
            out[f"{name}/{pen}"] = loss.detach().numpy()
            out[f"{name}/{pen}_dx"] = x.grad.numpy()
    x = t(gi.weights(91, (2, 2, 64, 80), 1.0)).requires_grad_()
    loss = smooothing_loss(x)
    loss.backward()
    out["smooth/loss"] = loss.detach().numpy(); out["smooth/dx"] = x.grad.numpy()

    l1 = REGISTRATIONModel.calculate_L1_loss
    a = t(gi.image(101, 2, (64, 64))).requires_grad_()
    b = t(gi.image(102, 2, (64, 64))).requires_grad_()
    mask = (b > -0.95) + (a > -0.95)
    loss = l1(None, a, b, mask)
    loss.backward()
    out["l1/loss"] = loss.detach().numpy(); out["l1/da"] = a.grad.numpy(); out["l1/db"] = b.grad.numpy()
    out["l1/mask_sum"] = mask.sum().numpy()
    out["l1/empty"] = np.array(float(l1(None, a.detach(), b.detach(), torch.zeros_like(mask))))
    out["l1/nomask"] = l1(None, a.detach(), b.detach(), None).numpy()
    save("losses", **out)


def det_randperm(counter):
    """Deterministic stand-in for torch.randperm (PatchSampleF draws, models/networks.py:609): the k-th
    call returns numpy RandomState(9000 + k).permutation(n).  Tests install the same function."""
    def f(n, device=None, **kw):
        counter[0] += 1
        return torch.from_numpy(np.random.RandomState(9000 + counter[0]).permutation(int(n))).to(device or 'cpu')
    return f


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in sd.items() if not k.endswith('.grid')}


def gen_nets():
    import_reference()
    from models import networks as rn
    from models.patchnce import PatchNCELoss
    import models.voxelmorph.torchvoxelmorph as rvxm
    import argparse
    out = {}

    # ---- layer-level: blur-pool down / up, instance norm (+relu), reflection pad: fwd + bwd
    torch.manual_seed(7)
    x = t(gi.weights(201, (2, 6, 12, 16), 1.0)).requires_grad_()
    for name, mod in (("down", rn.Downsample(6)), ("up", rn.Upsample(6))):
        y = mod(x)
        gy = t(gi.weights(202, tuple(y.shape), 1.0))
        y.backward(gy)
        out[f"layer/{name}"] = y.detach().numpy(); out[f"layer/{name}_dx"] = x.grad.numpy().copy(); x.grad = None
    blk = rn.ResnetBlock(6, 'reflect', rn.get_norm_layer('instance'), False, True)
    rn.init_weights(blk, 'xavier', 0.5)
    y = blk(x)
    y.backward(t(gi.weights(203, tuple(y.shape), 1.0)))
    out["layer/resblock"] = y.detach().numpy(); out["layer/resblock_dx"] = x.grad.numpy().copy()
    out.update(sd_np(blk.state_dict(), "layer/resblock_sd"))
    out.update({f"layer/resblock_grad/{k}": v.grad.numpy() for k, v in blk.named_parameters()})

    # ---- ResnetGenerator (ngf 8, 3 blocks, 32x40 input): full forward, encode_only features, backward
    torch.manual_seed(11)
    G = rn.define_G(1, 1, 8, 'resnet_4blocks', 'instance', False, 'xavier', 0.5, False, False, [], None)
    xin = t(gi.image_textured(211, 2, (32, 40))).requires_grad_()
    fake, feats = G(xin, [0, 4, 8, 12, 14], encode_only=False)
    out["G/in"] = xin.detach().numpy()
    out["G/fake"] = fake.detach().numpy()
    for i, f in enumerate(feats):
        out[f"G/feat{i}"] = f.detach().numpy()
    enc = G(xin, [0, 4, 8, 12, 14], encode_only=True)
    assert all(torch.equal(a, b) for a, b in zip(enc, feats))
    loss = (fake * t(gi.weights(212, tuple(fake.shape), 1.0))).sum() + sum(
        (f * t(gi.weights(213 + i, tuple(f.shape), 0.1))).sum() for i, f in enumerate(feats))
    loss.backward()
    out["G/dx"] = xin.grad.numpy()
    out.update(sd_np(G.state_dict(), "G/sd"))
    out.update({f"G/grad/{k}": v.grad.numpy() for k, v in G.named_parameters()})

    # ---- PatchSampleF + PatchNCELoss
    torch.manual_seed(12)
    opt = argparse.Namespace(netF_nc=32, batch_size=2, nce_T=0.07, nce_includes_all_negatives_from_minibatch=False)
    netF = rn.define_F(1, 'mlp_sample', 'instance', False, 'xavier', 0.5, False, [], opt)
    fq = [t(gi.weights(221, (2, 1, 20, 24), 1.0)).requires_grad_(), t(gi.weights(222, (2, 16, 10, 12), 1.0)).requires_grad_()]
    fk = [t(gi.weights(223, (2, 1, 20, 24), 1.0)), t(gi.weights(224, (2, 16, 10, 12), 1.0))]
    cnt = [0]
    torch_randperm = torch.randperm
    torch.randperm = det_randperm(cnt)
    try:
        k_pool, ids = netF(fk, 48, None)
    finally:
        torch.randperm = torch_randperm
    q_pool, _ = netF(fq, 48, ids)
    crit = PatchNCELoss(opt)
    total = 0
    for i, (q, k) in enumerate(zip(q_pool, k_pool)):
        l = crit(q, k)
        out[f"F/loss{i}"] = l.detach().numpy(); out[f"F/q{i}"] = q.detach().numpy(); out[f"F/k{i}"] = k.detach().numpy()
        out[f"F/ids{i}"] = ids[i].numpy()
        total = total + l.mean()
    total.backward()
    out["F/dq0"] = fq[0].grad.numpy(); out["F/dq1"] = fq[1].grad.numpy()
    out.update(sd_np(netF.state_dict(), "F/sd"))
    out.update({f"F/grad/{k}": v.grad.numpy() for k, v in netF.named_parameters()})

    # ---- VxmDense 2-D (shipped features, 64x64) and 3-D (small features, 16x16x16), bidir, fwd + bwd
    for name, shape, feats_ in (("R2", (64, 64), [[16, 32, 32, 64, 64, 64], [64, 64, 64, 32, 32, 32, 16]]),
                                ("R3", (16, 16, 16), [[8, 16, 16], [16, 16, 16, 8, 8]])):
        torch.manual_seed(13)
        R = rvxm.networks.VxmDense(shape, feats_, int_steps=7, bidir=True)
        with torch.no_grad():   # a visible deformation: the reference initialises the flow head at 1e-5
            R.flow.weight.mul_(2e4)
        src = t(gi.image_textured(231, 2, shape)).requires_grad_()
        tgt = t(gi.image_textured(232, 2, shape))
        ys, yt, flow = R(src, tgt)
        out[f"{name}/y_source"] = ys.detach().numpy(); out[f"{name}/y_target"] = yt.detach().numpy()
        out[f"{name}/pos_flow"] = flow.detach().numpy()
        ys2, flow2 = R(src, tgt, registration=True)
        assert torch.equal(flow2, flow)
        loss = (ys * t(gi.weights(233, tuple(ys.shape), 1.0))).sum() + (yt * t(gi.weights(234, tuple(yt.shape), 1.0))).sum() \
            + (flow * t(gi.weights(235, tuple(flow.shape), 0.1))).sum()
        loss.backward()
        out[f"{name}/d_src"] = src.grad.numpy()
        out.update(sd_np(R.state_dict(), f"{name}/sd"))
        out.update({f"{name}/grad/{k}": v.grad.numpy() for k, v in R.named_parameters()})
    save("nets", **out)


def gen_nets3d():
    """VxmDense 3-D with the reference's DEFAULT U-Net features (vxm/networks.py:9-14; the network of BASELINE
    configs[3]) on a 16 x 32 x 16 volume pair, not bidirectional: forward (registration=True) + backward."""
    import_reference()
    import models.voxelmorph.torchvoxelmorph as rvxm
    out = {}
    shape = (16, 32, 16)
    torch.manual_seed(17)
    R = rvxm.networks.VxmDense(shape, int_steps=7, bidir=False)
    with torch.no_grad():   # a visible deformation: the reference initialises the flow head at 1e-5
        R.flow.weight.mul_(2e4)
    src = t(gi.image_textured(241, 1, shape)).requires_grad_()
    tgt = t(gi.image_textured(242, 1, shape))
    ys, flow = R(src, tgt, registration=True)
    out["R3d/y_source"] = ys.detach().numpy(); out["R3d/pos_flow"] = flow.detach().numpy()
    loss = (ys * t(gi.weights(243, tuple(ys.shape), 1.0))).sum() + (flow * t(gi.weights(245, tuple(flow.shape), 0.1))).sum()
    loss.backward()
    out["R3d/d_src"] = src.grad.numpy()
    out.update(sd_np(R.state_dict(), "R3d/sd"))
    out.update({f"R3d/grad/{k}": v.grad.numpy() for k, v in R.named_parameters()})
    save("nets3d", **out)


def gen_step():
    """One full REGISTRATIONModel.optimize_parameters on CPU (reference code, B = 2, 64x64, ngf 8)."""
    import_reference()
    import shutil
    import tempfile
    work = tempfile.mkdtemp(prefix="dfmir_golden_")
    shutil.copy(os.path.join(REF, "deform256.jpg"), work)
    os.chdir(work)
    from options.train_options import TrainOptions
    from models import create_model
    import models.registration_model as rm
    B, S = 2, 64
    opt = TrainOptions(cmd_line=f'--dataroot x --name golden --CUT_mode CUT --no_flip --gpu_ids -1 --batch_size {B} '
                                f'--ngf 8 --crop_size {S} --load_size {S} --netF_nc 32 --num_patches 64 '
                                f'--checkpoints_dir {work}/ckpt').parse()
    # shim: the reference warps a batch-1 test image with a batch-B flow (registration_model.py:148-149),
    # which F.grid_sample rejects for B > 1; hand it a batch-B image of the crop size instead.
    dvf_img = t(gi.image_textured(301, B, (S, S), C=3))
    rm.open_image_to_torch = lambda path, size: dvf_img
    cnt = [0]
    torch_randperm = torch.randperm
    torch.randperm = det_randperm(cnt)
    try:
        torch.manual_seed(21)
        m = create_model(opt)
        data = {'A': t(gi.image_textured(302, B, (S, S))), 'B': t(gi.image_textured(303, B, (S, S))),
                'A_paths': [], 'B_paths': []}
        m.data_dependent_initialize(data)
        with torch.no_grad():
            m.netR.flow.weight.mul_(2e4)
        m.setup(opt)
        out = {}
        for n in ('G', 'F', 'R'):
            out.update(sd_np(getattr(m, 'net' + n).state_dict(), f"sd0/{n}"))
        cnt[0] = 100      # the step's randperm draws are calls 101..115
        m.set_input(data)
        m.optimize_parameters()
    finally:
        torch.randperm = torch_randperm
    for k, v in m.get_current_losses().items():
        out[f"loss/{k}"] = np.float32(v)
    for n in ('fake_B', 'idt_B', 'registered', 'regA', 'dvf'):
        out[f"vis/{n}"] = getattr(m, n).detach().numpy()
    for n in ('G', 'F', 'R'):
        net = getattr(m, 'net' + n)
        out.update({f"grad/{n}/{k}": v.grad.numpy() for k, v in net.named_parameters()})
        out.update(sd_np(net.state_dict(), f"sd1/{n}"))
    out["dvf_img"] = dvf_img.numpy()
    os.chdir(ROOT)
    save("step", **out)
    shutil.rmtree(work, ignore_errors=True)


SECTIONS = {"ops": gen_ops, "nets": gen_nets, "nets3d": gen_nets3d, "step": gen_step}

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    todo = sys.argv[1:] or list(SECTIONS)
    for s in todo:
        print(f"[gen_golden] {s}")
        SECTIONS[s]()
