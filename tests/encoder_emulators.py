"""Plain-torch emulations of the encoder kernel wrappers of b200lp/kernels.py (csrc/encoder.cu, csrc/mobilenet.cu) —
TEST INFRASTRUCTURE ONLY.  Each function restates the contract of one C-ABI entry point (include/b200lp.h) in whatever
dtype its inputs have; monkeypatched over `b200lp.kernels` they let the "not gpu" suite run the identity / pose encoder
SCHEDULES (embedders/resnext_native.py, embedders/mobilenet_native.py) in float64 against torchvision's own modules and
autograd.  On the GPU the same functions serve as the per-kernel references of tools/gpu_diag.py.
"""
import torch
import torch.nn.functional as F

STEM_KP = 192


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _split(y):
    return torch.stack([y, torch.zeros_like(y)])


def _unsplit(x):
    return x[0] + x[1] if x.dim() == 5 else x


def _act(v, act):
    if act == 1:
        return v.clamp_min(0)
    if act == 2:
        return v.clamp(0, 6)
    return v


def col_stats(x2d):
    return torch.stack([x2d.sum(0), (x2d * x2d).sum(0)])[None]


def bn_finalize(bn, part, count, training, want_stats=False):
    if training:
        s1, s2 = part[:, 0].sum(0), part[:, 1].sum(0)
        mean = s1 / count
        var = (s2 / count - mean * mean).clamp_min(0)
        if bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
                bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * count / max(count - 1, 1))
                bn.num_batches_tracked += 1
    else:
        mean, var = bn.running_mean.clone(), bn.running_var.clone()
    rstd = 1.0 / torch.sqrt(var + bn.eps)
    scale = bn.weight.detach() * rstd
    shift = bn.bias.detach() - mean * scale
    return (scale, shift, mean, rstd) if want_stats else (scale, shift)


def bn_act(x, scale=None, shift=None, res=None, res_scale=None, res_shift=None, act=1, round_tf32=True, want_f32=True,
           want_split=False, want_mask=False):
    v = x if scale is None else x * scale + shift
    if res is not None:
        v = v + (res if res_scale is None else res * res_scale + res_shift)
    v = _act(v, act)
    out = (v, _split(v)) if (want_f32 and want_split) else (v if want_f32 else _split(v))
    if want_mask:       # the emulated mask keeps one value per element (+1 / -1): mask_mode 4 tests `> 0`
        mask = torch.where(v > 0, torch.ones_like(v), -torch.ones_like(v))
        return (out + (mask,)) if isinstance(out, tuple) else (out, mask)
    return out


def _mask(dy, mask_mode, mask_src, x_raw, scale, shift):
    if mask_mode == 0:
        return dy
    if mask_mode in (1, 4):
        return dy * (mask_src > 0)
    pre = x_raw * scale + shift
    if mask_mode == 2:
        return dy * (pre > 0)
    return dy * ((pre > 0) & (pre < 6))


def bn_bwd(dy, x_raw, mean, rstd, gamma, scale=None, shift=None, mask_src=None, mask_mode=0, dgamma=None, dbeta=None,
           accumulate=False, batch_stats=True, round_tf32=False, want_dz=False):
    c = dy.shape[-1]
    dz = _mask(dy, mask_mode, mask_src, x_raw, scale, shift)
    xhat = (x_raw - mean) * rstd
    s1 = dz.reshape(-1, c).sum(0)
    s2 = (dz * xhat).reshape(-1, c).sum(0)
    m = dz.numel() // c
    if dgamma is None:
        dgamma, dbeta = s2.clone(), s1.clone()
    elif accumulate:
        dgamma.add_(s2); dbeta.add_(s1)
    else:
        dgamma.copy_(s2); dbeta.copy_(s1)
    if batch_stats:
        dx = gamma * rstd * (dz - s1 / m - xhat * (s2 / m))
    else:
        dx = gamma * rstd * dz
    return dx, dgamma, dbeta, (dz.clone() if want_dz else None)


def _gact(x, in_scale, in_shift):
    return x if in_scale is None else (x * in_scale + in_shift).clamp_min(0)


def gconv3x3_fwd(x, w, in_scale=None, in_shift=None, stride=1, want_stats=False):
    groups = x.shape[-1] // w.shape[1]
    y = _nhwc(F.conv2d(_nchw(_gact(x, in_scale, in_shift)), w.to(x.dtype), stride=stride, padding=1, groups=groups))
    return (y, col_stats(y.reshape(-1, y.shape[-1]))) if want_stats else y


def gconv_tensor_cores(n, h, w, c, cpg):
    return (h & (h - 1)) == 0 and (w & (w - 1)) == 0 and h >= 2 and w >= 2 and c % 32 == 0


def pack_gconv_weight(w, transpose=False, precision=0, out=None):
    return w            # the emulated data-gradient reads the OIHW weight itself


def zero_stuff2(x):
    n, h, w, c = x.shape
    out = torch.zeros((n, 2 * h, 2 * w, c), dtype=x.dtype, device=x.device)
    out[:, ::2, ::2] = x
    return out


def gconv3x3_wgrad_tc(x, dy, cpg, acc_into=None):
    return gconv3x3_wgrad(x, dy, cpg, acc_into=acc_into)


def gconv3x3_dgrad(dy, w, in_hw, stride=1, packed=None):
    n, ho, wo, c = dy.shape
    groups = c // w.shape[1]
    dx = torch.nn.grad.conv2d_input((n, c, in_hw[0], in_hw[1]), w.to(dy.dtype), _nchw(dy), stride=stride, padding=1,
                                    groups=groups)
    return _nhwc(dx)


def gconv3x3_wgrad(x, dy, cpg, in_scale=None, in_shift=None, stride=1, acc_into=None):
    c = x.shape[-1]
    g = torch.nn.grad.conv2d_weight(_nchw(_gact(x, in_scale, in_shift)), (c, cpg, 3, 3), _nchw(dy), stride=stride,
                                    padding=1, groups=c // cpg)
    if acc_into is not None:
        acc_into.add_(g)
        return acc_into
    return g.contiguous()


def im2col7x7_s2(x_nchw, want_f32=True, want_split=True):
    n, c, h, w = x_nchw.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    cols = F.unfold(x_nchw, kernel_size=7, padding=3, stride=2)            # (n, 147, ho*wo), row = c*49 + kh*7 + kw
    col = torch.zeros((n, ho, wo, STEM_KP), dtype=x_nchw.dtype, device=x_nchw.device)
    col[..., :147] = cols.transpose(1, 2).reshape(n, ho, wo, 147)
    if want_f32 and want_split:
        return col, _split(col)
    return col if want_f32 else _split(col)


def maxpool3x3s2_fwd(x, scale, shift, want_f32=True, want_split=False, want_idx=True, round_tf32=True):
    a = (x * scale + shift).clamp_min(0)
    n, h, w, c = a.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    pad = F.pad(_nchw(a), (1, 1, 1, 1), value=-1.0)
    win = pad.unfold(2, 3, 2).unfold(3, 3, 2).reshape(n, c, ho, wo, 9)      # taps kh*3 + kw
    y, idx = win.max(dim=-1)
    # first maximum like the kernel's strict '>' scan (torch.max returns an arbitrary one among ties on some builds)
    first = (win == y.unsqueeze(-1)).to(torch.uint8).argmax(dim=-1)
    y = _nhwc(y)
    return (y if want_f32 else None, _split(y) if want_split else None,
            _nhwc(first).to(torch.uint8) if want_idx else None)


def maxpool3x3s2_bwd(dy, idx, in_hw):
    n, ho, wo, c = dy.shape
    h, w = in_hw
    dx = torch.zeros((n, h + 2, w + 2, c), dtype=dy.dtype, device=dy.device)     # padded by 1
    for tap in range(9):
        kh, kw = tap // 3, tap % 3
        contrib = dy * (idx == tap)
        dx[:, kh:kh + 2 * ho:2, kw:kw + 2 * wo:2, :] += contrib
    return dx[:, 1:h + 1, 1:w + 1, :].contiguous()


def subsample2(x=None, x_split=None):
    y = x[:, ::2, ::2, :].contiguous() if x is not None else None
    ys = x_split[:, :, ::2, ::2, :].contiguous() if x_split is not None else None
    return y, ys


def scatter_add2(dsub, dx):
    dx[:, ::2, ::2, :] += dsub
    return dx


def avgpool_fwd(x):
    return x.mean((1, 2))


def avgpool_bwd(dy, hw):
    n, c = dy.shape
    return (dy / (hw[0] * hw[1]))[:, None, None, :].expand(n, hw[0], hw[1], c).contiguous()


def sgemm(a, b, trans_a=False, trans_b=False, acc_into=None, alpha=None, bias=None):
    r = (a.t() if trans_a else a) @ (b.t() if trans_b else b)
    if alpha is not None:
        r = r * alpha.reshape(()).to(r.dtype)
    if bias is not None:
        r = r + bias
    if acc_into is not None:
        acc_into.add_(r.view_as(acc_into))
        return acc_into
    return r.contiguous()


def pw_conv(x2d, weight, in_scale=None, in_shift=None, in_relu6=False, bias=None, want_stats=False):
    a = x2d if in_scale is None else x2d * in_scale + in_shift
    if in_scale is not None and in_relu6:
        a = a.clamp(0, 6)
    y = a @ weight.reshape(weight.shape[0], -1).t()
    part = col_stats(y) if want_stats else None
    if bias is not None:
        y = y + bias
    return (y, part) if want_stats else y


def mbv2_stem(x_nchw, weight, want_stats=False):
    y = _nhwc(F.conv2d(x_nchw, weight.to(x_nchw.dtype), stride=2, padding=1))
    return (y, col_stats(y.reshape(-1, 32))) if want_stats else y


def dw_conv3x3(x, weight, in_scale, in_shift, stride, want_stats=False):
    a = _nchw((x * in_scale + in_shift).clamp(0, 6))
    y = _nhwc(F.conv2d(a, weight.to(x.dtype), stride=stride, padding=1, groups=x.shape[-1]))
    return (y, col_stats(y.reshape(-1, y.shape[-1]))) if want_stats else y


def bn_apply(x, scale, shift, residual=None, relu6=False):
    y = x * scale + shift
    if relu6:
        y = y.clamp(0, 6)
    return y + residual if residual is not None else y


def bn_relu6_avgpool(x, scale, shift):
    return (x * scale + shift).clamp(0, 6).mean((1, 2))


def transpose2d(src):
    return src.t().contiguous()


def pw_wgrad(dy2d, x2d, in_scale=None, in_shift=None, in_relu6=False, acc_into=None):
    a = x2d if in_scale is None else x2d * in_scale + in_shift
    if in_scale is not None and in_relu6:
        a = a.clamp(0, 6)
    g = dy2d.t() @ a
    if acc_into is not None:
        acc_into.add_(g.view_as(acc_into))
        return acc_into
    return g.contiguous()


def dw_dgrad(dy, w, in_hw, stride):
    n, ho, wo, c = dy.shape
    dx = torch.nn.grad.conv2d_input((n, c, in_hw[0], in_hw[1]), w.to(dy.dtype), _nchw(dy), stride=stride, padding=1, groups=c)
    return _nhwc(dx)


def dw_wgrad(x, dy, in_scale, in_shift, stride, acc_into=None):
    c = x.shape[-1]
    a = _nchw((x * in_scale + in_shift).clamp(0, 6))
    g = torch.nn.grad.conv2d_weight(a, (c, 1, 3, 3), _nchw(dy), stride=stride, padding=1, groups=c)
    if acc_into is not None:
        acc_into.add_(g.view_as(acc_into))
        return acc_into
    return g.contiguous()


def mbv2_stem_wgrad(x_nchw, dy, acc_into=None):
    g = torch.nn.grad.conv2d_weight(x_nchw, (32, 3, 3, 3), _nchw(dy), stride=2, padding=1)
    if acc_into is not None:
        acc_into.add_(g.view_as(acc_into))
        return acc_into
    return g.contiguous()


def bias_grad(dy, acc_into=None):
    g = dy.reshape(-1, dy.shape[-1]).sum(0)
    if acc_into is not None:
        acc_into.add_(g)
        return acc_into
    return g


EMULATED = ["mbv2_stem", "dw_conv3x3", "bn_apply", "bn_relu6_avgpool", "transpose2d", "pw_wgrad", "dw_dgrad", "dw_wgrad",
            "mbv2_stem_wgrad", "bias_grad", "col_stats", "bn_finalize", "bn_act", "bn_bwd", "gconv3x3_fwd", "gconv3x3_dgrad", "gconv3x3_wgrad",
            "gconv_tensor_cores", "pack_gconv_weight", "zero_stuff2", "gconv3x3_wgrad_tc",
            "im2col7x7_s2", "maxpool3x3s2_fwd", "maxpool3x3s2_bwd", "subsample2", "scatter_add2", "avgpool_fwd",
            "avgpool_bwd", "sgemm", "pw_conv"]


def install(monkeypatch):
    import sys
    from b200lp import kernels as K
    this = sys.modules[__name__]
    for name in EMULATED:
        assert hasattr(K, name), name
        monkeypatch.setattr(K, name, getattr(this, name))
