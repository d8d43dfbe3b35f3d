"""Host-side helpers of codes/SRN/utils/util.py that the hot path and its callers use."""
import logging
import math
import os
import random
from datetime import datetime

import numpy as np
import torch


def get_timestamp():
    return datetime.now().strftime('%y%m%d-%H%M%S')


def mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)


def mkdirs(paths):
    for p in ([paths] if isinstance(paths, str) else paths):
        mkdir(p)


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def setup_logger(logger_name, root, phase, level=logging.INFO, screen=False):
    lg = logging.getLogger(logger_name)
    fmt = logging.Formatter('%(asctime)s.%(msecs)03d - %(levelname)s: %(message)s', datefmt='%y-%m-%d %H:%M:%S')
    fh = logging.FileHandler(os.path.join(root, phase + '_{}.log'.format(get_timestamp())), mode='w')
    fh.setFormatter(fmt)
    lg.setLevel(level)
    lg.addHandler(fh)
    if screen:
        sh = logging.StreamHandler()
        sh.setFormatter(fmt)
        lg.addHandler(sh)


def b_split(batch, mask):
    """(samples with mask 0, samples with mask 1).  The reference builds both by per-sample unsqueeze+cat
    (utils/util.py:150-163); the training mask is always [0]*B + [1]*B (DASR_model.py:176-179), for which
    this returns views batch[:B], batch[B:] — any other mask goes through index_select."""
    m = [int(v) for v in mask]
    n0 = m.count(0)
    if m == [0] * n0 + [1] * (len(m) - n0):
        fake, real = batch[:n0], batch[n0:]
    else:
        idx = torch.as_tensor(m, device=batch.device)
        fake, real = batch[idx == 0], batch[idx == 1]
    return (fake if fake.shape[0] else []), (real if real.shape[0] else [])


def b_merge(real_data, fake_data, mask):
    return torch.cat([(fake_data if int(m) == 0 else real_data)[i:i + 1] for i, m in enumerate(mask)])


def tensor2img(tensor, out_type=np.uint8, min_max=(0, 1)):
    """4D/3D/2D RGB tensor -> HWC BGR (or HW) numpy image in [0,255]; rounds for uint8.
    NOTE: like the reference, the input tensor is clamped IN PLACE when it is a CPU float tensor."""
    tensor = tensor.squeeze().float().cpu().clamp_(*min_max)
    tensor = (tensor - min_max[0]) / (min_max[1] - min_max[0])
    n_dim = tensor.dim()
    if n_dim == 4:
        from torchvision.utils import make_grid
        img = make_grid(tensor, nrow=int(math.sqrt(len(tensor))), normalize=False).numpy()
        img = np.transpose(img[[2, 1, 0], :, :], (1, 2, 0))
    elif n_dim == 3:
        img = np.transpose(tensor.numpy()[[2, 1, 0], :, :], (1, 2, 0))
    elif n_dim == 2:
        img = tensor.numpy()
    else:
        raise TypeError('Only support 4D, 3D and 2D tensor. But received with dimension: {:d}'.format(n_dim))
    if out_type == np.uint8:
        img = (img * 255.0).round()
    return img.astype(out_type)


def save_img(img, img_path, mode='RGB'):
    import cv2
    cv2.imwrite(img_path, img)


def calculate_psnr(img1, img2):
    mse = np.mean((img1.astype(np.float64) - img2.astype(np.float64)) ** 2)
    return float('inf') if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def ssim(img1, img2):
    import cv2
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    k = cv2.getGaussianKernel(11, 1.5)
    win = np.outer(k, k.transpose())
    f = lambda im: cv2.filter2D(im, -1, win)[5:-5, 5:-5]
    mu1, mu2 = f(img1), f(img2)
    s1, s2, s12 = f(img1 ** 2) - mu1 ** 2, f(img2 ** 2) - mu2 ** 2, f(img1 * img2) - mu1 * mu2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))).mean()


def calculate_ssim(img1, img2):
    if img1.shape != img2.shape:
        raise ValueError('Input images must have the same dimensions.')
    if img1.ndim == 2:
        return ssim(img1, img2)
    if img1.ndim == 3:
        if img1.shape[2] == 3:
            return np.array([ssim(img1, img2) for _ in range(3)]).mean()   # sic: the reference averages 3 identical calls
        if img1.shape[2] == 1:
            return ssim(np.squeeze(img1), np.squeeze(img2))
    raise ValueError('Wrong input image dimensions.')


def forward_chop(img, scale, model, shave=20, min_size=160000):
    """4-quadrant tiling with `shave` px overlap (utils/util.py:87-147), recursive above min_size."""
    h, w = img.size()[-2:]
    top, bottom = slice(0, h // 2 + shave), slice(h - h // 2 - shave, h)
    left, right = slice(0, w // 2 + shave), slice(w - w // 2 - shave, w)
    chops = [img[..., top, left], img[..., top, right], img[..., bottom, left], img[..., bottom, right]]
    if h * w < 4 * min_size:
        # the four quadrants have the same shape: ONE batched forward on the device (the reference runs them one at a time
        # through P.data_parallel(model, x, range(1)), utils/util.py:104-113)
        n = chops[0].shape[0]
        y = model(torch.cat([c.contiguous() for c in chops], 0))
        outs = [y[i * n:(i + 1) * n] for i in range(4)]
    else:
        outs = [forward_chop(c, scale, model, shave=shave, min_size=min_size) for c in chops]
    h, w = scale * h, scale * w
    top, bottom = slice(0, h // 2), slice(h - h // 2, h)
    bottom_r = slice(h // 2 - h, None)
    left, right = slice(0, w // 2), slice(w - w // 2, w)
    right_r = slice(w // 2 - w, None)
    b, c = outs[0].size()[:-2]
    y = outs[0].new_empty(b, c, h, w)
    y[..., top, left] = outs[0][..., top, left]
    y[..., top, right] = outs[1][..., top, right_r]
    y[..., bottom, left] = outs[2][..., bottom_r, left]
    y[..., bottom, right] = outs[3][..., bottom_r, right_r]
    return y
