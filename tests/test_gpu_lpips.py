"""GPU parity of the LPIPS row: dasr_b200.lpips against the fixture the REFERENCE's PNetLin produced
(tests/golden/lpips_alex.pt, oracle/gen_golden_lpips.py) and against the CPU oracle on ragged sizes."""
import pytest
import torch

from oracle import lpips_oracle as LP
from oracle import srn_oracle as O

pytestmark = pytest.mark.gpu


def _net(sd, lins):
    from dasr_b200.lpips import PerceptualLoss
    full = dict(sd)
    for i, w in enumerate(lins):
        full['lin%d.model.1.weight' % i] = w
    return PerceptualLoss(lin_weights=full, trunk_weights=full).cuda()


def test_lpips_value_and_gradient_vs_reference_fixture(golden):
    g = golden('lpips_alex.pt')
    sd = O.synth_state_dict(LP.alex_shapes(), g['w_seed'], 1.0)
    net = _net(sd, g['lins'])
    pred = O.synth_image(g['shape'], g['pred_seed']).cuda().requires_grad_(True)
    target = O.synth_image(g['shape'], g['target_seed']).cuda()
    val = net(pred, target, normalize=True)
    assert val.shape == g['value'].shape
    assert float((val.cpu() - g['value']).abs().max() / g['value'].abs().max()) < 1e-4
    val.mean().backward()
    err = float((pred.grad.cpu() - g['dpred']).abs().max() / g['dpred'].abs().max())
    assert err < 1e-3, err
    # the [-1, 1] entry point (validation metric: PerceptualLoss.forward(fake, real) without normalize)
    val2 = net(2 * pred.detach() - 1, 2 * target - 1)
    assert float((val2.cpu() - g['value']).abs().max() / g['value'].abs().max()) < 1e-4


@pytest.mark.parametrize('shape', [(1, 3, 35, 47), (3, 3, 128, 128)])
def test_lpips_ragged_sizes_vs_oracle(golden, shape):
    g = golden('lpips_alex.pt')
    sd = O.synth_state_dict(LP.alex_shapes(), 91, 1.0)
    net = _net(sd, g['lins'])
    pred = O.synth_image(shape, 92)
    target = O.synth_image(shape, 93)
    pc = pred.clone().requires_grad_(True)
    ref = LP.lpips(pc, target, sd, g['lins'])
    ref.mean().backward()
    pg = pred.cuda().requires_grad_(True)
    val = net(pg, target.cuda(), normalize=True)
    val.mean().backward()
    assert float((val.cpu() - ref.detach()).abs().max() / ref.detach().abs().max()) < 1e-4
    assert float((pg.grad.cpu() - pc.grad).abs().max() / pc.grad.abs().max()) < 1e-3


def test_lpips_feature_criterion_in_dasr_model(golden):
    """feature_criterion: "LPIPS" (train_DASR.json:80, train_DASR_auto_reproduce_aim2019.json) runs a training step and its
    log value equals the oracle's LPIPS of the same SR / HR pair."""
    from dasr_b200.srn.models import create_model
    from helpers import make_opt, unwrap
    g = golden('lpips_alex.pt')
    opt = make_opt(True, 'DASR', 1, 'wavelet')
    opt['train']['feature_criterion'] = 'LPIPS'
    opt['train']['feature_weight'] = 1e-2
    model = create_model(opt)
    sd = O.synth_state_dict(LP.alex_shapes(), 95, 1.0)
    full = dict(sd)
    for i, w in enumerate(g['lins']):
        full['lin%d.model.1.weight' % i] = w
    model.cri_fea.loss.loss_network.net.load_state_dict(full, strict=False)
    sdG = O.synth_state_dict(O.rrdbnet_shapes(nb=1), 96, 0.3)
    unwrap(model.netG).load_state_dict(sdG)
    B, h, w = 2, 16, 16
    data = {'LR_real': O.synth_image((B, 3, h, w), 1), 'LR_fake': O.synth_image((B, 3, h, w), 2),
            'HR': O.synth_image((B, 3, 4 * h, 4 * w), 3), 'HR_unpair': O.synth_image((B, 3, 4 * h, 4 * w), 4),
            'fake_w': O.synth_image((B, 1, h, w), 5)}
    model.feed_data(data, True)
    model.optimize_parameters(1)
    log = model.get_current_log()
    sr = O.rrdbnet_forward(data['LR_fake'], sdG, 1)
    ref = float(LP.lpips(sr, data['HR'], sd, g['lins']).mean())
    assert abs(log['loss/l_g_fea'] - ref) < 1e-3 * max(1.0, abs(ref)), (log['loss/l_g_fea'], ref)
    assert all(torch.isfinite(p.grad).all() for p in unwrap(model.netG).parameters() if p.grad is not None)
