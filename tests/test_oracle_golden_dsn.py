"""Pins oracle/dsn_oracle.py against fixtures produced by the imported reference DSN modules
(oracle/gen_golden_dsn.py -> tests/golden/dsn_*.pt).  CPU only."""
import torch

from oracle import dsn_oracle as D
from oracle import srn_oracle as O


def rel_linf(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_grads(p, g, tol=1e-4):
    for k, ref in g['grads'].items():
        assert rel_linf(p[k].grad, ref) < tol, k
    for k, n in g['grad_norms'].items():
        assert abs(float(p[k].grad.double().norm()) - n) <= tol * max(n, 1e-12), k


def test_de_resnet_forward_backward(golden):
    g = golden('dsn_de_resnet.pt')
    sd = D.synth_de_resnet(g['nres'], g['scale'], g['w_seed'], g['gain'])
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = O.synth_image(g['x_shape'], g['x_seed']).requires_grad_(True)
    out = D.de_resnet_forward(x, p, g['nres'], g['scale'])
    assert out.shape == g['out'].shape
    assert rel_linf(out.detach(), g['out']) < 1e-5
    (out * O.synth(tuple(out.shape), g['pat_seed'])).sum().backward()
    assert rel_linf(x.grad, g['dx']) < 1e-4
    check_grads(p, g)


def test_fs_discriminator(golden):
    g = golden('dsn_fsd.pt')
    for ft, n_in in (('wavelet', 9), ('gau', 3)):
        sd = O.synth_state_dict(D.fsd_shapes(n_in), g['w_seed'], 1.0)
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x = O.synth_image(g['x_shape'], g['x_seed']).requires_grad_(True)
        out = D.fsd_forward(x, p, None, ft)
        assert out.shape == g[ft]['out'].shape
        assert rel_linf(out.detach(), g[ft]['out']) < 1e-5, ft
        (out * O.synth(tuple(out.shape), g['pat_seed'])).sum().backward()
        assert rel_linf(x.grad, g[ft]['dx']) < 1e-4, ft
        check_grads(p, g[ft])


def test_generator_and_discriminator_losses(golden):
    g = golden('dsn_losses.pt')
    sdV = O.synth_state_dict(D.vgg16_shapes(), g['v_seed'], 1.0)
    tex = O.synth_image((2, 1, 16, 16), g['tex_seed']).requires_grad_(True)
    out = O.synth_image((2, 3, 32, 32), g['out_seed']).requires_grad_(True)
    tgt = O.synth_image((2, 3, 32, 32), g['tgt_seed'])
    total, parts = D.g_loss(tex, out, tgt, sdV)
    total.backward()
    for k, ref in (('tex', 'tex_loss'), ('per', 'per_loss'), ('col', 'col_loss')):
        assert abs(float(parts[k]) - float(g[ref])) <= 1e-5 * abs(float(g[ref])), k
    assert abs(float(total) - float(g['total'])) <= 1e-5 * abs(float(g['total']))
    assert rel_linf(tex.grad, g['dtex']) < 1e-4 and rel_linf(out.grad, g['dout']) < 1e-4
    real = O.synth_image((2, 1, 16, 16), g['real_seed']).requires_grad_(True)
    fake = O.synth_image((2, 1, 16, 16), g['fake_seed']).requires_grad_(True)
    dl = D.discriminator_loss(real, fake)
    dl.backward()
    assert abs(float(dl) - float(g['d_loss'])) <= 1e-5 * abs(float(g['d_loss']))
    assert rel_linf(real.grad, g['dreal']) < 1e-5 and rel_linf(fake.grad, g['dfake']) < 1e-5


def test_dsn_train_iterations(golden):
    g = golden('dsn_step.pt')
    s = g['seeds']
    sdG = D.synth_de_resnet(g['nres'], g['scale'], s['G'], g['gain_G'])
    sdD = O.synth_state_dict(D.fsd_shapes(9), s['D'], 1.0)
    sdV = O.synth_state_dict(D.vgg16_shapes(), s['V'], 1.0)
    optG = O.AdamState(sdG, 1e-4, 0.5)
    optD = O.AdamState(sdD, 1e-4, 0.5)
    for it in range(g['steps']):
        inp = O.synth_image((2, 3, 128, 128), s['inp'] + it)
        bic = O.synth_image((2, 3, 32, 32), s['bic'] + it)
        dis = O.synth_image((2, 3, 32, 32), s['dis'] + it)
        log, gG, gD, fake = D.dsn_train_step(sdG, sdD, sdV, inp, bic, dis, optG, optD, g['nres'], g['scale'])
        for k, v in g['logs'][it].items():
            assert abs(log[k] - v) <= 2e-4 * max(abs(v), 1e-6), (it, k, log[k], v)
        if it == 0:
            assert rel_linf(fake, g['first']['fake']) < 1e-5
            for k, n in g['first']['gnG'].items():
                assert abs(float(gG[k].double().norm()) - n) <= 1e-3 * max(n, 1e-12), k
            for k, n in g['first']['gnD'].items():
                assert abs(float(gD[k].double().norm()) - n) <= 1e-3 * max(n, 1e-12), k
    for k, ref in g['paramsG'].items():
        assert rel_linf(sdG[k], ref) < 1e-4, k
    for k, ref in g['paramsD'].items():
        assert rel_linf(sdD[k], ref) < 1e-4, k


def test_lpips_alex_oracle(golden):
    """oracle/lpips_oracle.py (prepared for the LPIPS feature criterion) against the reference's PNetLin."""
    from oracle import lpips_oracle as LP
    g = golden('lpips_alex.pt')
    sd = O.synth_state_dict(LP.alex_shapes(), g['w_seed'], 1.0)
    pred = O.synth_image(g['shape'], g['pred_seed']).requires_grad_(True)
    target = O.synth_image(g['shape'], g['target_seed'])
    val = LP.lpips(pred, target, sd, g['lins'])
    assert val.shape == g['value'].shape
    assert rel_linf(val.detach(), g['value']) < 1e-5
    val.mean().backward()
    assert rel_linf(pred.grad, g['dpred']) < 1e-4
