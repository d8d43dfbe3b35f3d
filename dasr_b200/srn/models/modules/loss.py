"""Losses of the SRN hot path on dasr_b200 reduction kernels (reference: codes/SRN/models/modules/loss.py
and the inline loss arithmetic of DASR_model.py:210-284).  Every loss is one fused
value+gradient kernel pair (deterministic two-stage reduction) wrapped in an autograd.Function."""
import torch
import torch.nn as nn

from dasr_b200 import ops


class _L1Function(torch.autograd.Function):
    """mean(w * |a - b|)  (w: [N,1,H,W] broadcast over channels, or None).  Gradient flows to `a` only."""

    @staticmethod
    def forward(ctx, a, b, w):
        a4 = a if a.dim() == 4 else a.reshape(1, 1, 1, -1)
        b4 = b if b.dim() == 4 else b.reshape(1, 1, 1, -1)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        grad = torch.empty_like(a4, dtype=torch.float32) if a.requires_grad else None
        ops.wl1_loss(a4.contiguous().float(), b4.contiguous().float(), w.contiguous().float() if w is not None else None,
                     loss, grad, 1.0)
        ctx.grad = grad
        ctx.shape = a.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None
        return (ctx.grad * g).reshape(ctx.shape), None, None


class _MSEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        grad = torch.empty_like(a, dtype=torch.float32) if a.requires_grad else None
        ops.mse_loss(a.contiguous().float(), b.contiguous().float(), loss, grad, 1.0)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x, dtype=torch.float32) if x.requires_grad else None
        ops.bce_logits_loss(x.contiguous().float(), target, loss, grad, 1.0)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None


class _MeanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        out = torch.empty((), dtype=torch.float32, device=x.device)
        ops.mean(x.contiguous().float(), out)
        ctx.shape, ctx.n = x.shape, x.numel()
        return out

    @staticmethod
    def backward(ctx, g):
        return (g / ctx.n).expand(ctx.shape)


def weighted_l1(a, b, w=None):
    return _L1Function.apply(a, b, w)


def mean(x):
    return _MeanFunction.apply(x)


class L1Loss(nn.Module):
    """nn.L1Loss() (mean reduction) on the fused kernel."""

    def forward(self, a, b):
        return _L1Function.apply(a, b, None)


class MSELoss(nn.Module):
    def forward(self, a, b):
        return _MSEFunction.apply(a, b)


class GANLoss(nn.Module):
    """[vanilla | lsgan | wgan-gp] with constant real/fake labels (loss.py:8-40)."""

    def __init__(self, gan_type, real_label_val=1.0, fake_label_val=0.0):
        super().__init__()
        self.gan_type = gan_type.lower()
        self.real_label_val, self.fake_label_val = real_label_val, fake_label_val
        if self.gan_type not in ('vanilla', 'lsgan', 'wgan-gp'):
            raise NotImplementedError('GAN type [{:s}] is not found'.format(self.gan_type))

    def forward(self, input, target_is_real):
        if self.gan_type == 'wgan-gp':
            m = mean(input)
            return -1 * m if target_is_real else m
        t = self.real_label_val if target_is_real else self.fake_label_val
        if self.gan_type == 'vanilla':
            return _BCEFunction.apply(input, t)
        return _MSEFunction.apply(input, torch.full_like(input, t))


class _HaarFunction(torch.autograd.Function):
    """J=1 Haar split + DASR normalisation + band-major concat (DASR_model.py:442-452) in one kernel."""

    @staticmethod
    def forward(ctx, x, norm):
        N, C, H, W = x.shape
        ll = torch.empty((N, C, H // 2, W // 2), dtype=torch.float32, device=x.device)
        hc = torch.empty((N, 3 * C, H // 2, W // 2), dtype=torch.float32, device=x.device)
        ops.haar_fwd(x.contiguous().float(), ll, hc, norm)
        ctx.cfg = (norm, x.shape)
        return ll, hc

    @staticmethod
    def backward(ctx, dll, dhc):
        norm, shape = ctx.cfg
        dx = torch.empty(shape, dtype=torch.float32, device=dll.device)
        ops.haar_bwd(dll.contiguous().float(), dhc.contiguous().float(), dx, norm)
        return dx, None


def haar_split(x, norm):
    return _HaarFunction.apply(x, bool(norm))
