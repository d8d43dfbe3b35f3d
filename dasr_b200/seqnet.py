"""Layer-list executor for the small sequential CNNs of the DSN path (De_resnet, DiscriminatorBasic) on the fp32
implicit-GEMM conv kernels.  NHWC fp32 activations; every op is a dasr_b200 C-ABI kernel (no torch operator computes).

A net is a list of layer dicts:
  {'op': 'conv', 'k', 's', 'p', 'w': param index, 'b': param index | None, 'act': ACT_NONE | ACT_LRELU}
  {'op': 'prelu', 'a': param index}            nn.PReLU() with one slope
  {'op': 'in_lrelu'}                           InstanceNorm2d(affine=False, eps=1e-5) + LeakyReLU(0.2), in place
  {'op': 'bn_lrelu', 'w', 'b': param indices, 'rm', 'rv', 'nbt': running_mean / running_var / num_batches_tracked buffers,
   'training': bool}                           BatchNorm2d(affine, eps=1e-5, momentum=0.1) + LeakyReLU(0.2)
  {'op': 'sigmoid'}                            in place
  {'op': 'pixel_shuffle', 'r': 2}              nn.PixelShuffle(r)
  conv layers may carry 'view': (cout, cin, k, k) — an nn.Linear weight [out, in] used as a k x k valid conv over the
  NHWC feature map it flattens (Linear(512*4*4, 100) after .view(N, -1) == conv k=4 with the weight viewed [100,512,4,4])
  {'op': 'res_begin'} ... {'op': 'res_end'}    y = x_at_begin + y
"""
import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, DGRAD


def _out_hw(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def forward(x, layers, params, save=True):
    """x: NCHW fp32 CUDA tensor.  Returns (NCHW output, ctx)."""
    if not x.is_cuda:
        raise ops._lib.DasrError('dasr_b200 DSN networks need CUDA tensors; there is no CPU fallback')
    N, C0, H, W = x.shape
    a = torch.empty((N, H, W, C0), dtype=torch.float32, device=x.device)
    ops.nchw_to_nhwc(x.contiguous().float(), a)
    last, ctx = forward_nhwc(a, layers, params, save)
    out = torch.empty((last.shape[0], last.shape[3], last.shape[1], last.shape[2]), dtype=torch.float32, device=x.device)
    ops.nhwc_to_nchw(last, out)
    if ctx is not None:
        ctx['shape'] = (N, C0, H, W)
    return out, ctx


def forward_nhwc(a, layers, params, save=True):
    """a: NHWC fp32 activation.  Returns (NHWC fp32 output, ctx)."""
    x = a
    H, W = a.shape[1], a.shape[2]
    acts, aux, res_stack = [a], [], []
    for L in layers:
        cur = acts[-1]
        op = L['op']
        if op == 'conv':
            wt = params[L['w']]
            if L.get('view') is not None:
                wt = wt.view(*L['view'])
            bs = params[L['b']] if L['b'] is not None else None
            n, h, w, _ = cur.shape
            oh, ow = _out_hw(h, L['k'], L['s'], L['p']), _out_hw(w, L['k'], L['s'], L['p'])
            if oh <= 0 or ow <= 0:
                raise ops._lib.DasrError('input %dx%d too small for the network' % (H, W))
            o = torch.empty((n, oh, ow, wt.shape[0]), dtype=torch.float32, device=x.device)
            ops.conv2d_f32(cur, ops.pack_filter_f32(wt), bs, o, L['k'], L['s'], L['p'], act=L.get('act', ACT_NONE), slope=0.2)
            aux.append(None)
        elif op == 'prelu':
            o = torch.empty_like(cur)
            ops.prelu_fwd(cur, params[L['a']], o)
            aux.append(None)
        elif op == 'in_lrelu':
            st = torch.empty((cur.shape[0], cur.shape[3], 2), dtype=torch.float32, device=x.device)
            ops.instnorm_lrelu_fwd(cur, st, 1e-5, 0.2)
            o = cur
            aux.append(st)
        elif op == 'bn_lrelu':
            st = torch.empty((cur.shape[3], 2), dtype=torch.float32, device=x.device)
            o = torch.empty_like(cur)
            ops.bn_lrelu_fwd(cur, o, params[L['w']], params[L['b']], L['rm'], L['rv'], st, 1e-5, 0.1, L['training'], 0.2)
            if L['training'] and L.get('nbt') is not None:
                L['nbt'].add_(1)
            aux.append(st)
        elif op == 'sigmoid':
            ops.sigmoid_fwd(cur, cur)
            o = cur
            aux.append(None)
        elif op == 'pixel_shuffle':
            r = L['r']
            n, h, w, cc = cur.shape
            o = torch.empty((n, h * r, w * r, cc // (r * r)), dtype=torch.float32, device=x.device)
            ops.pixel_shuffle(cur, o, r)
            aux.append(None)
        elif op == 'res_begin':
            res_stack.append(cur)
            o = cur
            aux.append(None)
        elif op == 'res_end':
            skip = res_stack.pop()
            ops.axpby(cur, 1.0, skip, 1.0, cur)
            o = cur
            aux.append(None)
        else:
            raise ValueError(op)
        acts.append(o)
    ctx = dict(acts=acts, aux=aux) if save else None
    return acts[-1], ctx


def backward(ctx, layers, params, dout, need_dx=True, need_dw=True):
    """Returns (dx NCHW | None, grads list aligned with params (None where a param got no gradient))."""
    acts = ctx['acts']
    N, C0, H, W = ctx['shape']
    g = torch.empty(tuple(acts[-1].shape), dtype=torch.float32, device=dout.device)
    ops.nchw_to_nhwc(dout.contiguous().float(), g)
    g, grads = backward_nhwc(ctx, layers, params, g, need_dx, need_dw)
    dx = None
    if need_dx:
        dx = torch.empty((N, C0, H, W), dtype=torch.float32, device=dout.device)
        ops.nhwc_to_nchw(g, dx)
    return dx, grads


def backward_nhwc(ctx, layers, params, g, need_dx=True, need_dw=True):
    """g: NHWC fp32 gradient of the output (may be modified in place).  Returns (NHWC input gradient | None, grads)."""
    acts, aux = ctx['acts'], ctx['aux']
    dout = g
    grads = [None] * len(params)
    skip_stack = []
    first_conv = next(i for i, L in enumerate(layers) if L['op'] == 'conv')
    for li in reversed(range(len(layers))):
        L = layers[li]
        op = L['op']
        if op == 'conv':
            if L.get('act', ACT_NONE) == ACT_LRELU:
                if skip_stack and skip_stack[-1] is g:
                    g = g.clone()
                ops.act_bwd(g, acts[li + 1], 0.2)
            wt = params[L['w']]
            if L.get('view') is not None:
                wt = wt.view(*L['view'])
            act_kind = L.get('act', ACT_NONE)
            if act_kind == ops.ACT_RELU:
                if skip_stack and skip_stack[-1] is g:
                    g = g.clone()
                ops.act_bwd(g, acts[li + 1], 0.0)
            if need_dw:
                gw = torch.empty_like(wt, dtype=torch.float32)
                grads[L['w']] = gw.view_as(params[L['w']])
                db = None
                if L['b'] is not None:
                    db = grads[L['b']] = torch.empty_like(params[L['b']], dtype=torch.float32)
                ops.conv2d_wgrad_f32(acts[li], g, gw, db, L['k'], L['s'], L['p'])
            if li > first_conv or need_dx:
                gin = torch.empty(tuple(acts[li].shape), dtype=torch.float32, device=dout.device)
                ops.conv2d_f32(g, ops.pack_filter_f32(wt, for_dgrad=True), None, gin, L['k'], L['s'], L['p'], mode=DGRAD)
                g = gin
        elif op == 'prelu':
            gz = torch.empty_like(g)
            da = torch.empty_like(params[L['a']], dtype=torch.float32)
            ops.prelu_bwd(acts[li], g, params[L['a']], gz, da)
            if need_dw:
                grads[L['a']] = da
            g = gz
        elif op == 'in_lrelu':
            gz = torch.empty_like(g)
            ops.instnorm_lrelu_bwd(acts[li + 1], aux[li], g, gz, 0.2)
            g = gz
        elif op == 'bn_lrelu':
            gz = torch.empty_like(g)
            dgm = torch.empty_like(params[L['w']], dtype=torch.float32) if need_dw else None
            dbt = torch.empty_like(params[L['b']], dtype=torch.float32) if need_dw else None
            ops.bn_lrelu_bwd(acts[li], acts[li + 1], g, params[L['w']], aux[li], gz, dgm, dbt, L['training'], 0.2)
            if need_dw:
                grads[L['w']], grads[L['b']] = dgm, dbt
            g = gz
        elif op == 'sigmoid':
            gz = torch.empty_like(g)
            ops.sigmoid_bwd(acts[li + 1], g, gz)
            g = gz
        elif op == 'pixel_shuffle':
            gz = torch.empty(tuple(acts[li].shape), dtype=torch.float32, device=dout.device)
            ops.pixel_shuffle(g, gz, L['r'], inverse=True)
            g = gz
        elif op == 'res_end':
            skip_stack.append(g)            # the same gradient feeds the skip connection and the residual branch
        elif op == 'res_begin':
            gs = skip_stack.pop()
            ops.axpby(g, 1.0, gs, 1.0, g)
    return (g if need_dx else None), grads


class SeqFunction(torch.autograd.Function):
    """autograd node of one layer-list network; params in state_dict order."""

    @staticmethod
    def forward(ctx, x, layers, *params):
        ctx.need_dx = x.requires_grad
        ctx.need_dw = any(p.requires_grad for p in params)
        out, saved = forward(x, layers, [p.detach() for p in params], save=ctx.need_dx or ctx.need_dw)
        ctx.saved, ctx.params, ctx.layers = saved, params, layers
        return out

    @staticmethod
    def backward(ctx, dout):
        # ctx.saved is kept: the DSN iteration back-propagates through D(fake) twice (D loss, then G loss)
        dx, grads = backward(ctx.saved, ctx.layers, [p.detach() for p in ctx.params], dout, ctx.need_dx, ctx.need_dw)
        return (dx, None) + tuple(grads)
