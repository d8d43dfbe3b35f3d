"""create_model(opt) — same dispatch on opt['model'] as codes/SRN/models/__init__.py:5-26 for the models
of the SRN hot path.  'DASR_FS_ESRGAN_patchGAN' (the name the shipped train_DASR*.json files use, which
the reference's create_model does not know — SURVEY.md §5.6a) is accepted as an alias of 'DASR'."""
import logging

logger = logging.getLogger('base')


def create_model(opt):
    model = opt['model']
    if model == 'sr':
        from .SR_model import SRModel as M
    elif model in ('DASR', 'DASR_FS_ESRGAN_patchGAN'):
        from .DASR_model import DASR_Model as M
    else:
        raise NotImplementedError('Model [{:s}] not recognized.'.format(str(model)))
    m = M(opt)
    logger.info('Model [{:s}] is created.'.format(m.__class__.__name__))
    return m
