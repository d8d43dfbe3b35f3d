"""Functional stand-in for pytorch_wavelets (not installed, not vendored by the reference, version
unpinned).  Only J=1 'haar' on even H, W — the one configuration DASR uses (DASR_model.py:56).
Follows the package's documented output convention: (Yl, [Yh]) with Yh[:, :, 0..2] = LH, HL, HH and
pywt 'haar' analysis filters.  Used ONLY by oracle/gen_golden.py to let the reference code run."""
import torch
import torch.nn as nn


class DWTForward(nn.Module):
    def __init__(self, J=1, wave='haar', mode='zero'):
        super().__init__()
        assert J == 1 and wave == 'haar'

    def forward(self, x):
        assert x.shape[-1] % 2 == 0 and x.shape[-2] % 2 == 0
        a = x[:, :, 0::2, 0::2]
        b = x[:, :, 0::2, 1::2]
        c = x[:, :, 1::2, 0::2]
        d = x[:, :, 1::2, 1::2]
        ll = (a + b + c + d) * 0.5
        lh = (a + b - c - d) * 0.5
        hl = (a - b + c - d) * 0.5
        hh = (a - b - c + d) * 0.5
        return ll, [torch.stack((lh, hl, hh), dim=2)]

    def cuda(self, device=None):      # DSN/loss.py:103 calls .cuda() unconditionally; the stand-in has no state
        return self


class DWTInverse(nn.Module):
    def __init__(self, wave='haar', mode='zero'):
        super().__init__()
