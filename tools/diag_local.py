"""Diagnostic: is d(local NCE)/d(regA) sensitive to the VALUE of regA (intrinsic) or is the engine wrong?"""
import contextlib, io, os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import inputs as gi
from oracle import torch_port as tp
from dfmir_b200 import registration_model as rm
import dfmir_b200.functional as Fn

S = 256
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sds = tp.random_state_dicts(ngf=64, n_blocks=9, crop=S, seed=5)
sds[2]['flow.weight'] = sds[2]['flow.weight'] * 2e4
A = torch.from_numpy(gi.image_textured(700 + B, B, (S, S)))
Bm = torch.from_numpy(gi.image_textured(710 + B, B, (S, S)))
SIZES = [(S + 6) ** 2, S * S, (S // 2) ** 2, (S // 4) ** 2, (S // 4) ** 2]


def ids_for(call):
    return [torch.from_numpy(np.random.RandomState(9000 + 100 + call * 5 + i + 1).permutation(n))[:256].cuda() for i, n in enumerate(SIZES)]


P = {n: {k: v.double().cuda().clone().requires_grad_(v.is_floating_point() and not k.endswith(('.filt', '.grid'))) for k, v in sd.items()}
     for n, sd in zip('GFR', sds)}
Ad, Bd = A.double().cuda(), Bm.double().cuda()
regA_o, _, pos_flow_o = tp.vxm_dense(Ad, Bd, P['R'], 6, 7)


def o_local(regA, layers=(0, 1, 2, 3, 4)):
    L = [0, 4, 8, 12, 16]
    fq = tp.resnet_generator(regA, P['G'], 9, L, encode_only=True)
    fk = tp.resnet_generator(Bd, P['G'], 9, L, encode_only=True)
    ids = ids_for(2)
    k_pool, _ = tp.patch_sample(fk, P['F'], 256, ids)
    q_pool, _ = tp.patch_sample(fq, P['F'], 256, ids)
    tot = 0.0
    per = []
    for i, (q, k) in enumerate(zip(q_pool, k_pool)):
        l = (tp.patchnce(q, k, B) * 0.25).mean() / 5 * 0.25
        per.append(l)
        if i in layers:
            tot = tot + l
    return tot, per


def o_grad(regA_val):
    x = regA_val.detach().double().clone().requires_grad_()
    t, per = o_local(x)
    g = torch.autograd.grad(t, x, retain_graph=True)[0]
    gl = [torch.autograd.grad(p, x, retain_graph=True)[0] for p in per]
    return float(t), g, gl


def rel(a, b):
    a, b = a.double(), b.double()
    return f"relerr {float((a - b).norm() / b.norm()):.3e} cos {float((a * b).sum() / (a.norm() * b.norm())):.5f} |a|/|b| {float(a.norm() / b.norm()):.4f}"


def ours(engineR):
    opt = rm.default_options(batch_size=B, crop_size=S, load_size=S, gpu_ids=[0])
    with contextlib.redirect_stdout(io.StringIO()):
        m = rm.REGISTRATIONModel(opt)
        m.data_dependent_initialize({'A': A, 'B': Bm})
        m.setup(opt)
    for n, sd in zip('GFR', sds):
        getattr(m, 'net' + n).load_state_dict(sd, strict=False)
    m.set_input({'A': A, 'B': Bm})
    m.forward()
    Fn.CONV_ENGINE = engineR
    y = m.netR(m.real_A, m.real_B)
    Fn.CONV_ENGINE = "auto"
    return m, y[0]


def ours_grad(m, regA_val, engineG="auto"):
    x = regA_val.detach().float().clone().requires_grad_()
    Fn.CONV_ENGINE = engineG
    t = m.calculate_NCE_loss(m.real_B, x, [i for i in ids_for(2)]) * 0.25
    g = torch.autograd.grad(t, x)[0]
    Fn.CONV_ENGINE = "auto"
    return float(t), g


t0, g0, gl0 = o_grad(regA_o)
print("oracle at oracle regA: loss", t0, "|g|", float(g0.norm()), "per-layer |g|", [f"{float(x.norm()):.3e}" for x in gl0])
for engR in ("auto", "simt"):
    m, regA = ours(engR)
    print(f"=== R engine {engR}: regA ours vs oracle", rel(regA, regA_o), "max abs", float((regA.double() - regA_o).abs().max()))
    t1, g1, gl1 = o_grad(regA)
    print("  oracle G at OUR regA vs oracle G at oracle regA (intrinsic sensitivity):", rel(g1, g0), "loss", t1)
    for i, (a, b) in enumerate(zip(gl1, gl0)):
        print(f"     layer {i}:", rel(a, b))
    for engG in ("auto", "simt"):
        t2, g2 = ours_grad(m, regA, engG)
        print(f"  our G[{engG}] at our regA vs oracle G at our regA:   ", rel(g2, g1), "loss", t2)
        print(f"  our G[{engG}] at our regA vs oracle G at oracle regA:", rel(g2, g0))
        t3, g3 = ours_grad(m, regA_o, engG)
        print(f"  our G[{engG}] at oracle regA vs oracle G at oracle regA:", rel(g3, g0), "loss", t3)
