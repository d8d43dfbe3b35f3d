import sys, torch
sys.path.insert(0, '.')
from oracle import srn_oracle as O
from dasr_b200.srn.models.modules.architecture import VGGFeatureExtractor
sdfull = O.synth_state_dict(O.vgg19_shapes(34), 7, 1.0)
for fl in (1, 3, 4, 6, 8, 9, 11, 13, 15, 17, 18, 20, 26, 27, 29, 34):
    net = VGGFeatureExtractor(feature_layer=fl, weights=sdfull).cuda()
    res = {}
    for prec in ('fp32', 'bf16'):
        net.precision = prec
        x = O.synth_image((2, 3, 64, 48), 3).cuda().requires_grad_(True)
        out = net(x)
        (out * O.synth(tuple(out.shape), 5).cuda()).sum().backward()
        res[prec] = (out.detach().float(), x.grad.clone())
    ef = float((res['bf16'][0] - res['fp32'][0]).norm() / res['fp32'][0].norm())
    eg = float((res['bf16'][1] - res['fp32'][1]).norm() / res['fp32'][1].norm())
    print('feature_layer %2d  out %s  feat rel-L2 %.3e   dx rel-L2 %.3e' % (fl, tuple(out.shape), ef, eg))
