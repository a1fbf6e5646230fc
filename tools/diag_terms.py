"""Diagnostic: gradient of every loss term of the step w.r.t. the generator output `fake`, ours vs the float64 oracle."""
import contextlib, io, os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from oracle import torch_port as tp
from dfmir_b200 import registration_model as rm, losses
import dfmir_b200.functional as Fn

S = 256
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sds = tp.random_state_dicts(ngf=64, n_blocks=9, crop=S, seed=5)
if os.environ.get('FLOWSCALE', '1') == '1':
    sds[2]['flow.weight'] = sds[2]['flow.weight'] * 2e4
A = torch.from_numpy(gi.image_textured(700 + B, B, (S, S)))
Bm = torch.from_numpy(gi.image_textured(710 + B, B, (S, S)))


def ids_for(call):            # 5 layers per NCE call
    sizes = [(S + 6) ** 2, S * S, (S // 2) ** 2, (S // 4) ** 2, (S // 4) ** 2]
    return [torch.from_numpy(np.random.RandomState(9000 + 100 + call * 5 + i + 1).permutation(n))[:256].cuda() for i, n in enumerate(sizes)]


# ---------------- oracle, float64
P = {n: {k: v.double().cuda().clone().requires_grad_(v.is_floating_point() and not k.endswith(('.filt', '.grid'))) for k, v in sd.items()}
     for n, sd in zip('GFR', sds)}
Ad, Bd = A.double().cuda(), Bm.double().cuda()
fake = tp.resnet_generator(torch.cat((Ad, Bd), 0), P['G'], 9)
fake_B, idt_B = fake[:B], fake[B:]
regA, _, pos_flow = tp.vxm_dense(Ad, Bd, P['R'], 6, 7)
registered = tp.spatial_transform(fake_B, pos_flow)


def o_nce(src, tgt, call):
    fq = tp.resnet_generator(tgt, P['G'], 9, [0, 4, 8, 12, 16], encode_only=True)
    fk = tp.resnet_generator(src, P['G'], 9, [0, 4, 8, 12, 16], encode_only=True)
    ids = ids_for(call)
    k_pool, _ = tp.patch_sample(fk, P['F'], 256, ids)
    q_pool, _ = tp.patch_sample(fq, P['F'], 256, ids)
    tot = 0.0
    for q, k in zip(q_pool, k_pool):
        tot = tot + (tp.patchnce(q, k, B) * 0.25).mean()
    return tot / 5


o_terms = {
    'NCE': o_nce(Ad, fake_B, 0), 'NCE_Y': o_nce(Bd, idt_B, 1), 'local': o_nce(Bd, regA, 2) * 0.25,
    'L1a': tp.masked_l1(registered, Bd, (Bd > -0.95) + (registered > -0.95)),
    'L1b': tp.masked_l1(idt_B, registered, (idt_B > -0.95) + (registered > -0.95)),
    'smooth': tp.smoothing(pos_flow) * 0.2,
}
o_g = {}
for k, t in o_terms.items():
    gs = torch.autograd.grad(t, [fake, pos_flow, P['G']['model.4.weight'], regA], retain_graph=True, allow_unused=True)
    o_g[k] = [None if g is None else g.detach().cpu() for g in gs]
    print("oracle", k, float(t))

# ---------------- ours, under several switches
def run_ours(tag):
    opt = rm.default_options(batch_size=B, crop_size=S, load_size=S, gpu_ids=[0])
    with contextlib.redirect_stdout(io.StringIO()):
        m = rm.REGISTRATIONModel(opt)
        m.data_dependent_initialize({'A': A, 'B': Bm})
        m.setup(opt)
    for n, sd in zip('GFR', sds):
        getattr(m, 'net' + n).load_state_dict(sd, strict=False)
    m.set_input({'A': A, 'B': Bm})
    m.forward()
    _eng = Fn.CONV_ENGINE
    if getattr(Fn, "_R_FP32", False):
        Fn.CONV_ENGINE = "simt"
    y = m.netR(m.real_A, m.real_B)
    Fn.CONV_ENGINE = _eng
    flow = y[2]
    m.registered = m.spatialTransformer(m.fake_B, flow)
    m.regA = y[0]
    dev = m.device
    terms = {
        'NCE': m.calculate_NCE_loss(m.real_A, m.fake_B, [t.to(dev) for t in ids_for(0)]),
        'NCE_Y': m.calculate_NCE_loss(m.real_B, m.idt_B, [t.to(dev) for t in ids_for(1)]),
        'local': m.calculate_NCE_loss(m.real_B, m.regA, [t.to(dev) for t in ids_for(2)]) * 0.25,
    }
    la, lb = m._masked_l1_pair((m.registered, m.real_B, m.real_B, m.registered), (m.idt_B, m.registered, m.idt_B, m.registered))
    terms['L1a'], terms['L1b'] = la, lb
    terms['smooth'] = rm.smooothing_loss(flow) * 0.2
    w4 = m.netG.model[4].weight
    for k, t in terms.items():
        gs = torch.autograd.grad(t, [m.fake, flow, w4, m.regA], retain_graph=True, allow_unused=True)
        line = f"[{tag}] {k:7s} ours {float(t):.6f} oracle {float(o_terms[k]):.6f}"
        for name, g, ref in zip(("d/dfake", "d/dflow", "d/dw4", "d/dregA"), gs, o_g[k]):
            if g is None or ref is None:
                continue
            g = g.detach().cpu().double()
            if g.is_sparse:
                g = g.to_dense()
            rel = float((g - ref).norm() / (ref.norm() + 1e-300))
            cos = float((g * ref).sum() / (g.norm() * ref.norm() + 1e-300))
            line += f" | {name} |ref| {float(ref.norm()):.3e} relerr {rel:.3e} cos {cos:.5f}"
        print(line)


for tag, setup in [("default", lambda: None),
                   ("no_stats_epilogue", lambda: setattr(Fn, "STATS_IN_EPILOGUE", False)),
                   ("convs_fp32_only", lambda: setattr(Fn, "UMMA_MIN_POSITIONS", 10 ** 9)),
                   ("R_fp32", lambda: setattr(Fn, "_R_FP32", True)),
                   ("simt", lambda: setattr(Fn, "CONV_ENGINE", "simt"))]:
    Fn.SPARSE_TAP_GRAD, Fn.STATS_IN_EPILOGUE, Fn.S2D_STRIDED, rm.REUSE_REAL_FEATURES, Fn.CONV_ENGINE = True, True, True, True, "auto"
    Fn.UMMA_MIN_POSITIONS = 4096
    Fn._R_FP32 = False
    setup()
    run_ours(tag)
