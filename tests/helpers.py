"""Shared helpers of the GPU tests (importable as `helpers`: tests/conftest.py puts this directory on sys.path)."""
import torch


def make_opt(is_train, model, nb=1, fs='wavelet', gpu=True):
    from dasr_b200.srn.options.options import dict_to_nonedict
    return dict_to_nonedict({
        'name': 'test', 'model': model, 'scale': 4, 'gpu_ids': [0] if gpu else None, 'is_train': is_train, 'chop': False,
        'val_lpips': False, 'multiweights': True,
        'path': {'pretrain_model_G': None, 'pretrain_model_D_target': None, 'pretrain_model_D_source': None,
                 'models': '/tmp', 'training_state': '/tmp'},
        'network_G': {'which_model_G': 'RRDB_net', 'norm_type': None, 'mode': 'CNA', 'nf': 64, 'nb': nb, 'in_nc': 3,
                      'out_nc': 3, 'gc': 32, 'group': 1, 'scale': 4},
        'network_D': {'which_model_D': 'discriminator_patch', 'which_model_pairD': 'discriminator_patch',
                      'norm_type': 'Batch', 'act_type': 'leakyrelu', 'mode': 'CNA', 'nf': 64,
                      'in_nc': 9 if fs == 'wavelet' else 3, 'n_layers': 2},
        'train': {'lr_G': 5e-5, 'weight_decay_G': 0, 'beta1_G': 0.9, 'lr_D': 5e-5, 'weight_decay_D': 0, 'beta1_D': 0.9,
                  'lr_scheme': 'MultiStepLR', 'lr_steps': [50000, 80000], 'lr_gamma': 0.5, 'fs': fs, 'norm': True,
                  'sup_LL': True, 'fs_kernel_size': 5, 'pixel_criterion': 'l1', 'pixel_weight': 1, 'pixel_LL_weight': 1,
                  'feature_criterion': 'l1', 'feature_weight': 1e-2, 'gan_type': 'vanilla', 'ragan': False,
                  'gan_H_target': 1e-4, 'gan_H_source': 0, 'G_update_inter': 1, 'D_update_inter': 1,
                  'D_update_ratio': 1, 'D_init_iters': 0, 'manual_seed': 0, 'niter': 10, 'val_freq': 10}})


def unwrap(net):
    return net.module if isinstance(net, torch.nn.DataParallel) else net


# --------------------------------------------------------------------------------------------------
# float64 "truth" evaluation of the mirror modules with plain torch ops (CPU), for ill-conditioned gradients:
# BatchNorm stacks amplify fp32 rounding differences between two correct implementations (torch CPU vs torch GPU already
# differ by percents on the BatchNorm FS discriminator), so a gradient passes when it is within tolerance of the
# reference fixture OR at least as close to the float64 result of the same algorithm as the reference's fp32 run is (x3).
# --------------------------------------------------------------------------------------------------

def native_forward(mod, x):
    """Evaluate a mirror module tree with torch's own operators (no dasr_b200 kernel), any dtype / device."""
    import torch.nn as nn
    import torch.nn.functional as F
    name = mod.__class__.__name__
    if isinstance(mod, nn.Conv2d):
        return F.conv2d(x, mod.weight, mod.bias, mod.stride, mod.padding)
    if name == 'ShortcutBlock':
        return x + native_forward(mod.sub, x)
    if name == 'ResNetBlock':
        return x + native_forward(mod.res, x) * mod.res_scale
    if isinstance(mod, nn.Sequential):
        for m in mod.children():
            x = native_forward(m, x)
        return x
    if name == 'SRResNet':
        return native_forward(mod.model, x)
    if name == 'Discriminator_VGG_128':
        lr = lambda t: F.leaky_relu(t, 0.2)
        f = lr(native_forward(mod.conv0_0, x))
        for n in ('0_1', '1_0', '1_1', '2_0', '2_1', '3_0', '3_1', '4_0', '4_1'):
            f = lr(getattr(mod, 'bn' + n)(native_forward(getattr(mod, 'conv' + n), f)))
        f = f.reshape(f.size(0), -1)
        return mod.linear2(lr(mod.linear1(f)))
    if name == 'Discriminator_VGG_192':
        f = native_forward(mod.features, x)
        return mod.classifier(f.reshape(f.size(0), -1))
    if name == 'DiscriminatorBasic':
        return native_forward(mod.net, x)
    return mod(x)


def truth64(net, x, pat, forward=None):
    """float64 CPU copy of `net` evaluated with torch ops: returns (out, dx, {param name: grad})."""
    import copy
    n64 = copy.deepcopy(net).cpu().double()
    n64.train(net.training)
    x64 = x.detach().cpu().double().requires_grad_(True)
    out = (forward or native_forward)(n64, x64)
    (out * pat.detach().cpu().double().reshape(out.shape)).sum().backward()
    return out.detach(), x64.grad, {k: p.grad for k, p in n64.named_parameters() if p.grad is not None}


def as_good_as_reference(got, ref, truth, tol=1e-3):
    """Acceptance of a gradient that passes through BatchNorm / LeakyReLU stacks:
      1. |got - ref| within tol (rel L-inf) of the reference fixture, or
      2. got at least as close to the float64 truth as the reference's fp32 run is (x3, rel L2), or
      3. isolated LeakyReLU kink flips: a pre-activation within fp32 rounding of zero takes the other slope in two correct
         fp32 implementations (among ~1e7 activations it happens to a few), which moves the gradients downstream of that one
         element by up to ~1e-2 of the maximum while everything else agrees to 1e-6.  tools/debug_f1b.py shows torch's own
         GPU operators, our kernels and the CPU reference each hit by it on different sub-chains
         (profiles/r2_kink_flips.txt).  Accepted: rel-L2 error vs truth < 5e-3 and rel-L-inf < 5e-2."""
    got, ref, truth = got.detach().double().cpu(), ref.detach().double().cpu(), truth.detach().double().cpu()
    if float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)) < tol:
        return True
    tn = truth.norm().clamp_min(1e-30)
    e2 = float((got - truth).norm() / tn)
    if e2 <= 3.0 * float((ref - truth).norm() / tn) + 1e-6:
        return True
    einf = float((got - truth).abs().max() / truth.abs().max().clamp_min(1e-30))
    return e2 < 5e-3 and einf < 5e-2
