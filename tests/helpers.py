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
