"""JSON option files of codes/SRN/options (``//`` comments allowed) -> nested dicts.

Same entry points as the reference (parse, NoneDict, dict_to_nonedict, dict2str, check_resume;
options/options.py:8-121) so train.py / test.py / Auto_Reproduce.py keep working, and more forgiving
where the shipped configs crash the reference: ``gpu_ids`` may be null/[]
"""
import json
import logging
import os
import os.path as osp
from collections import OrderedDict


def _strip_comments(path):
    with open(path, 'r') as f:
        return ''.join(line.split('//')[0] + '\n' for line in f)


def parse(opt_path, is_train=True):
    opt = json.loads(_strip_comments(opt_path), object_pairs_hook=OrderedDict)
    opt['is_train'] = is_train
    scale = opt['scale']

    for phase, ds in opt['datasets'].items():
        ds['phase'] = phase.split('_')[0]
        ds['scale'] = scale
        is_lmdb = False
        for key in ('dataroot_HR', 'dataroot_HR_bg', 'dataroot_LR'):
            if ds.get(key) is not None:
                ds[key] = os.path.expanduser(ds[key])
                if key != 'dataroot_HR_bg' and ds[key].endswith('lmdb'):
                    is_lmdb = True
        ds['data_type'] = 'lmdb' if is_lmdb else 'img'
        if ds['phase'] == 'train' and ds.get('subset_file') is not None:
            ds['subset_file'] = os.path.expanduser(ds['subset_file'])

    for key, path in opt['path'].items():
        if path:
            opt['path'][key] = os.path.expanduser(path)
    root = opt['path']['root']
    if is_train:
        exp = os.path.join(root, 'experiments', opt['name'])
        opt['path'].update(experiments_root=exp, models=os.path.join(exp, 'models'),
                           training_state=os.path.join(exp, 'training_state'), log=exp,
                           val_images=os.path.join(exp, 'val_images'))
        if 'debug' in opt['name']:
            opt['train']['val_freq'] = 8
            opt['logger']['print_freq'] = 2
            opt['logger']['save_checkpoint_freq'] = 8
            opt['train']['lr_decay_iter'] = 10
    else:
        res = os.path.join(root, 'results', opt['name'])
        opt['path'].update(results_root=res, log=res)

    opt['network_G']['scale'] = scale

    gpu_ids = opt.get('gpu_ids') or []
    gpu_list = ','.join(str(x) for x in gpu_ids)
    if gpu_list and 'LOCAL_RANK' not in os.environ:   # one-process-per-GPU launches pin devices themselves
        os.environ['CUDA_VISIBLE_DEVICES'] = gpu_list
        print('export CUDA_VISIBLE_DEVICES=' + gpu_list)
    return opt


class NoneDict(dict):
    def __missing__(self, key):
        return None


def dict_to_nonedict(opt):
    if isinstance(opt, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in opt.items()})
    if isinstance(opt, list):
        return [dict_to_nonedict(v) for v in opt]
    return opt


def dict2str(opt, indent_l=1):
    msg = ''
    for k, v in opt.items():
        pad = ' ' * (indent_l * 2)
        if isinstance(v, dict):
            msg += pad + k + ':[\n' + dict2str(v, indent_l + 1) + pad + ']\n'
        else:
            msg += pad + k + ': ' + str(v) + '\n'
    return msg


def check_resume(opt):
    logger = logging.getLogger('base')
    if opt['path']['resume_state']:
        if opt['path']['pretrain_model_G'] or opt['path']['pretrain_model_D']:
            logger.warning('pretrain_model path will be ignored when resuming training.')
        idx = osp.basename(opt['path']['resume_state']).split('.')[0]
        opt['path']['pretrain_model_G'] = osp.join(opt['path']['models'], '{}_G.pth'.format(idx))
        logger.info('Set [pretrain_model_G] to ' + opt['path']['pretrain_model_G'])
        if 'gan' in opt['model']:
            opt['path']['pretrain_model_D'] = osp.join(opt['path']['models'], '{}_D.pth'.format(idx))
            logger.info('Set [pretrain_model_D] to ' + opt['path']['pretrain_model_D'])
