"""Plain-torch emulations of the tensor-level kernel wrappers in b200lp/kernels.py — TEST INFRASTRUCTURE ONLY.

Each function restates the *contract* of one C-ABI entry point (include/b200lp.h) with torch CPU operators, in whatever
dtype its inputs have (the schedule tests run them in float64).  Monkeypatched over `b200lp.kernels` they let the
"not gpu" suite execute the plugins' host logic — which kernel is called on which tensor, layouts, residual modes,
spectral-norm bookkeeping, gradient sinks — and compare it with the oracle / the reference's golden vectors without a
GPU.  Nothing under latent-pose-reenactment_b200/ imports this file; the kernels themselves are checked against torch
on the GPU (tools/gpu_diag.py).
"""
import torch
import torch.nn.functional as F

TF32, BF16X3 = 0, 1


# ------------------------------------------------------------------------------------------------ weights, convs
def pack_conv_weight(w_oihw, scale=None, transpose=False, precision=TF32, out=None):
    co, ci, kh, kw = w_oihw.shape
    w = w_oihw.detach() * (scale if scale is not None else 1.0)
    if transpose:       # wp[ci][T-1-tap][co]
        p = w.reshape(co, ci, kh * kw).flip(2).permute(1, 2, 0).contiguous()
    else:               # wp[co][tap][ci]
        p = w.reshape(co, ci, kh * kw).permute(0, 2, 1).contiguous()
    if precision != TF32:
        p = torch.stack([p, torch.zeros_like(p)])          # (hi, lo) planes: no rounding in the emulation
    if out is not None:
        out.copy_(p)
        return out
    return p


def _unpack(wp, ksize):
    if wp.dim() == 4:
        wp = wp[0] + wp[1]
    rows, taps, cols = wp.shape
    assert taps == ksize * ksize
    return wp.reshape(rows, ksize, ksize, cols).permute(0, 3, 1, 2)       # -> (rows, cols, k, k) = conv2d weight


def conv_fwd(x, wp, ksize, bias=None, residual=None, residual_mode=0, relu=False, round_tf32=False, block_n=0,
             out=None, emit_split=False, stages=0, scale=None, ctas_per_sm=0, splits=0, variant=0, a_stages=0, grouped=0):
    if x.dim() == 5:
        x = x[0] + x[1]
    if grouped:         # the emulated grouped packing (encoder_emulators.pack_gconv_weight) is the OIHW weight itself
        w = wp.to(x.dtype)
        y = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1, groups=x.shape[-1] // grouped)
    else:
        w = _unpack(wp, ksize).to(x.dtype)
        y = F.conv2d(x.permute(0, 3, 1, 2), w, padding=ksize // 2)
    if scale is not None:
        y = y * scale.to(y.dtype)
    if bias is not None:
        y = y + bias.to(y.dtype)[None, :, None, None]
    if residual is not None and residual_mode == 1:
        y = y + residual.permute(0, 3, 1, 2)
    elif residual is not None and residual_mode == 2:
        y = y + F.interpolate(residual.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    if relu:
        y = y.relu()
    y = y.permute(0, 2, 3, 1).contiguous()
    if out is not None:
        out.copy_(y)
        y = out
    return (y, torch.stack([y, torch.zeros_like(y)])) if emit_split else y


def conv_wgrad(x, dy, ksize, scale=1.0, kstep=0, stages=0, splits=0):
    cin, cout = x.shape[3], dy.shape[3]
    g = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, ksize, ksize), dy.permute(0, 3, 1, 2),
                                    padding=ksize // 2)
    return (g * scale).contiguous()


def conv_wgrad_sn_acc(x, dy, ksize, grad, weight=None, inv_sigma=None, u=None, v=None, accumulate=True):
    g = conv_wgrad(x, dy, ksize)
    if inv_sigma is not None:
        s = inv_sigma.reshape(())
        dot = (g * weight.detach()).sum()
        g = s * g - s * s * dot * torch.outer(u, v).reshape(g.shape)
    if accumulate:
        grad.add_(g)
    else:
        grad.copy_(g)
    return grad


def bias_grad(dy, acc_into=None):
    db = dy.reshape(-1, dy.shape[-1]).sum(0)
    if acc_into is not None:
        acc_into.add_(db)
        return acc_into
    return db


# ------------------------------------------------------------------------------------------------ spectral norm
def sn_scratch(w):
    return torch.empty(1, dtype=w.dtype, device=w.device)


def sn_sigma_multi(layers, training):
    inv, snaps = [], []
    for (w, u, v, eps, _scratch) in layers:
        wm = w.reshape(w.shape[0], -1)
        if training:
            vn = torch.mv(wm.t(), u)
            vn = vn / vn.norm().clamp_min(eps)
            un = torch.mv(wm, vn)
            un = un / un.norm().clamp_min(eps)
            v.copy_(vn)
            u.copy_(un)
        inv.append(1.0 / torch.dot(u, torch.mv(wm, v)))
        snaps.append((u.clone(), v.clone()))
    return torch.stack(inv), snaps


def sn_wgrad_fix(g, w, inv_sigma, u, v):
    s = inv_sigma.reshape(())
    return s * g - s * s * (g * w.detach()).sum() * torch.outer(u, v).reshape(g.shape)


# ------------------------------------------------------------------------------------------------ AdaIN
def in_stats(x, eps):
    mean = x.mean((1, 2))
    var = x.var((1, 2), unbiased=False)
    return mean, (var + eps).rsqrt()


def _adain(x, mean, rstd, gamma, beta, upsample2):
    y = ((x - mean[:, None, None, :]) * rstd[:, None, None, :] * gamma[:, None, None, :] + beta[:, None, None, :]).relu()
    if upsample2:
        y = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
    return y


def adain_relu(x, mean, rstd, gamma, beta, upsample2=False, round_tf32=True, want_f32=True, want_split=False):
    y = _adain(x, mean, rstd, gamma, beta, upsample2).contiguous()
    ys = torch.stack([y, torch.zeros_like(y)]) if want_split else None
    if want_f32 and want_split:
        return y, ys
    return y if want_f32 else ys


def adain_stats_apply(x, gamma, beta, eps, upsample2=False, round_tf32=True, want_f32=True, want_split=False):
    mean, rstd = in_stats(x, eps)
    return mean, rstd, adain_relu(x, mean, rstd, gamma, beta, upsample2=upsample2, round_tf32=round_tf32, want_f32=want_f32,
                                  want_split=want_split)


def adain_relu_bwd(x, mean, rstd, gamma, beta, dy, upsample2=False, add=None, round_tf32=False):
    """Autograd through the emulated forward INCLUDING the statistics' dependence on x."""
    with torch.enable_grad():
        xr = x.detach().requires_grad_(True)
        gr = gamma.detach().clone().requires_grad_(True)
        br = beta.detach().clone().requires_grad_(True)
        m = xr.mean((1, 2))
        eps = (1.0 / rstd.detach() ** 2 - x.detach().var((1, 2), unbiased=False))      # recover the layer's eps
        r = (xr.var((1, 2), unbiased=False) + eps).rsqrt()
        y = _adain(xr, m, r, gr, br, upsample2)
        dx, dg, db = torch.autograd.grad(y, [xr, gr, br], dy)
    if add is not None:
        dx = dx + add
    return dx, dg, db


# ------------------------------------------------------------------------------------------------ elementwise
def nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def relu_round(x):
    return x.relu()


def relu_bwd(y, dy):
    return dy * (y > 0)


def relu_bwd_fused(y, dy, add=None, want_quarter=False, bias_a=None, bias_b=None, round_tf32=False):
    dx = (dy if add is None else dy + add) * (y > 0)
    col = dx.reshape(-1, dx.shape[-1]).sum(0)
    for b in (bias_a, bias_b):
        if b is not None:
            b.add_(col)
    return (dx, 0.25 * dx) if want_quarter else dx


def avgpool2(x, addend=None, round_tf32=False):
    y = F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()
    return y + addend if addend is not None else y


def avgpool2_bwd(dy):
    return (dy * 0.25).repeat_interleave(2, 1).repeat_interleave(2, 2).contiguous()


def upsample2_bwd(dy):
    n, h2, w2, c = dy.shape
    return dy.reshape(n, h2 // 2, 2, w2 // 2, 2, c).sum((2, 4))


def l1_sum(a, b, out, scale):
    out += scale * (a - b).abs().sum()


def l1_sum_code(a, b, out, scale):
    out += scale * (a - b).abs().sum()
    return torch.where(a > 0, torch.sign(a - b) + 2, torch.zeros_like(a))      # the emulated code keeps one value per element


def l1_code_bwd(code, shape, gscale, scale2, d_in=None):
    g = (gscale.reshape(()) if gscale is not None else 1.0) * scale2
    d = (code - 2) * g + (d_in if d_in is not None else 0)
    return torch.where(code > 0, d, torch.zeros_like(d)).reshape(shape)


def l1_sum_code_pool(a, b, out, scale, ap=None, bp=None):
    pa, pb = avgpool2(a), avgpool2(b)
    if ap is not None:
        ap.copy_(pa); pa = ap
    if bp is not None:
        bp.copy_(pb); pb = bp
    return l1_sum_code(a, b, out, scale), pa, pb


def l1_code_bwd_unpool(code, shape, gscale, scale2, d_low):
    return l1_code_bwd(code, shape, gscale, scale2, d_in=avgpool2_bwd(d_low))


def l1_bwd(a, b, gscale, scale2, da=None):
    g = torch.sign(a - b) * (gscale.reshape(()) * scale2)
    if da is not None:
        da.add_(g)
        return da
    return g


def l1_relu_bwd(a, b, gscale, scale2, d_in=None):
    g = torch.sign(a - b) * (gscale.reshape(()) * scale2)
    if d_in is not None:
        g = g + d_in
    return g * (a > 0)


# ------------------------------------------------------------------------------------------------ Cin = 3 stems
def conv3x3_c3_fwd(x_nchw, w, wscale=None, bias=None, pre_scale=None, pre_shift=None, relu=False, round_tf32=False,
                   tensor_cores=None):
    x = x_nchw
    if pre_scale is not None:
        x = x * pre_scale[None, :, None, None] + pre_shift[None, :, None, None]
    y = F.conv2d(x, w.detach() * (wscale.reshape(()) if wscale is not None else 1.0), bias, padding=1)
    if relu:
        y = y.relu()
    return y.permute(0, 2, 3, 1).contiguous()


def im2col3x3_c3(x_nchw):
    n, c, h, w = x_nchw.shape
    col = F.unfold(x_nchw, 3, padding=1).reshape(n, 27, h, w).permute(0, 2, 3, 1)      # column index = c*9 + kh*3 + kw
    return torch.cat([col, col.new_zeros(n, h, w, 5)], dim=3).contiguous()


def col2im3x3_c3(dcol, pre_scale=None):
    n, h, w, _ = dcol.shape
    dx = F.fold(dcol[..., :27].permute(0, 3, 1, 2).reshape(n, 27, h * w), (h, w), 3, padding=1)
    if pre_scale is not None:
        dx = dx * pre_scale[None, :, None, None]
    return dx.contiguous()


# ------------------------------------------------------------------------------------------------ generator tail
def _compose(t):
    segm = t[..., 3:] * 0.5 + 0.5
    rgb = (t[..., :3] * 0.75 + 0.5) * segm
    return rgb.permute(0, 3, 1, 2).contiguous(), segm.permute(0, 3, 1, 2).contiguous()


def gen_tail_fwd(x, w, wscale, bias):
    a = F.conv2d(x.permute(0, 3, 1, 2), w.detach() * wscale.reshape(()), bias, padding=1).permute(0, 2, 3, 1)
    t = torch.tanh(a).contiguous()
    rgbs, segm = _compose(t)
    return rgbs, segm, t


def gen_tail_compose(a32, bias):
    t = torch.tanh(a32[..., :4] + bias).contiguous()
    rgbs, segm = _compose(t)
    return rgbs, segm, t


def gen_tail_bwd_act(t, d_rgbs, d_segm, stride=32):
    with torch.enable_grad():
        a = torch.atanh(t.detach().clamp(-1 + 1e-15, 1 - 1e-15)).requires_grad_(True)
        rgbs, segm = _compose(torch.tanh(a))
        outs, gs = [], []
        if d_rgbs is not None:
            outs.append(rgbs); gs.append(d_rgbs)
        if d_segm is not None:
            outs.append(segm); gs.append(d_segm)
        (da,) = torch.autograd.grad(outs, [a], gs)
    if stride > 4:
        da = torch.cat([da, da.new_zeros(da.shape[:3] + (stride - 4,))], dim=3)
    return da.contiguous()


def gen_tail_bwd(x, t, w, wscale, d_rgbs, d_segm, need_dx=True, need_dw=True):
    da = gen_tail_bwd_act(t, d_rgbs, d_segm, stride=4)
    dx = dw = db = None
    if need_dx:
        dx = F.conv_transpose2d(da.permute(0, 3, 1, 2), w.detach() * wscale.reshape(()), padding=1).permute(0, 2, 3, 1)
        dx = dx.contiguous()
    if need_dw:
        dw = conv_wgrad(x, da, 3)
        db = bias_grad(da)
    return dx, dw, db


# ------------------------------------------------------------------------------------------------ buffer copies
def copy_plan(pairs):
    return dict(pairs=list(pairs), sig=None, count=len(pairs))


def copy_multi(plan):
    for d, s in plan["pairs"]:
        d.copy_(s)


# ------------------------------------------------------------------------------------------------ small dense / losses
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


def dice_fwd(fake, real, weight):
    s0, s1, s2 = (fake * real).sum(), (fake * fake).sum(), (real * real).sum()
    sums = torch.stack([s0, s1, s2])
    return (-torch.log(2 * s0 / (s1 + s2)) * weight).reshape(1), sums


def dice_bwd(fake, real, sums, grad, weight):
    rs = real.sum(1, keepdim=True)
    return -weight * grad.reshape(()) * (rs / sums[0] - 2 * fake / (sums[1] + sums[2]))


def adversarial_fwd(fake_g, fake_d, real, relativistic=0):
    assert relativistic == 0
    return torch.stack([-fake_g.mean(), torch.relu(1 - real).mean() + torch.relu(1 + fake_d).mean()])


def adversarial_bwd(fake_d, real, grad_g, grad_d, need_g=True, need_d=True):
    b = real.numel()
    dg = dd = dr = None
    if need_g and grad_g is not None:
        dg = torch.full_like(real, -1.0 / b) * grad_g.reshape(())
    if need_d and grad_d is not None:
        dd = (1 + fake_d > 0).to(real.dtype) * grad_d.reshape(()) / b
        dr = -(1 - real > 0).to(real.dtype) * grad_d.reshape(()) / b
    return dg, dd, dr


def _crop_grid(boxes, shape):
    t, b, l, r = boxes.t()
    n, c, h, w = shape
    theta = torch.zeros(n, 2, 3, dtype=boxes.dtype, device=boxes.device)
    theta[:, 0, 0] = (r - l) / w
    theta[:, 1, 1] = (b - t) / h
    theta[:, 0, 2] = (l + r) / w - 1
    theta[:, 1, 2] = (t + b) / h - 1
    return F.affine_grid(theta, list(shape), align_corners=False)


def crop_bilinear_fwd(x, boxes, out_hw=None):
    shape = tuple(x.shape[:2]) + tuple(out_hw or x.shape[2:])
    return F.grid_sample(x, _crop_grid(boxes.to(x.dtype), shape), mode="bilinear", padding_mode="reflection", align_corners=False)


def crop_bilinear_bwd(dy, boxes, in_hw):
    x = torch.zeros(tuple(dy.shape[:2]) + tuple(in_hw), dtype=dy.dtype, device=dy.device, requires_grad=True)
    with torch.enable_grad():
        y = F.grid_sample(x, _crop_grid(boxes.to(dy.dtype), tuple(dy.shape)), mode="bilinear", padding_mode="reflection",
                          align_corners=False)
        (g,) = torch.autograd.grad(y, x, dy)
    return g


def disc_head_fwd(feat, embed, w, inv_sigma, bias):
    o = feat.clamp_min(0).sum((1, 2))
    score = (o @ w) * inv_sigma.reshape(()) + bias.reshape(())
    if embed is not None:
        score = score + (o * embed).sum(1)
    return score, o


def disc_head_bwd(feat, embed, w, inv_sigma, o, grad, need_feat=True, need_embed=True, need_params=True, dw_acc=None,
                  db_acc=None):
    s = inv_sigma.reshape(())
    d_feat = d_embed = dw = ds = db = None
    if need_feat:
        k = grad[:, None] * (s * w[None, :] + (embed if embed is not None else 0))
        d_feat = (feat > 0).to(feat.dtype) * k[:, None, None, :]
        if need_embed and embed is not None:
            d_embed = grad[:, None] * o
    if need_params:
        a = (grad[:, None] * o).sum(0)
        ds = (a * w).sum().reshape(1)
        if dw_acc is not None:
            dw_acc.add_(s * a); dw = dw_acc
            db_acc.add_(grad.sum()); db = db_acc
        else:
            dw, db = s * a, grad.sum().reshape(1)
    return d_feat, d_embed, dw, ds, db


EMULATED = [
    "sgemm", "dice_fwd", "dice_bwd", "adversarial_fwd", "adversarial_bwd", "crop_bilinear_fwd", "crop_bilinear_bwd",
    "disc_head_fwd", "disc_head_bwd",
    "pack_conv_weight", "conv_fwd", "conv_wgrad", "conv_wgrad_sn_acc", "bias_grad", "sn_scratch", "sn_sigma_multi",
    "sn_wgrad_fix", "in_stats", "adain_relu", "adain_stats_apply", "adain_relu_bwd", "nchw_to_nhwc", "nhwc_to_nchw", "relu_round", "relu_bwd", "relu_bwd_fused",
    "avgpool2", "avgpool2_bwd", "upsample2_bwd", "l1_sum", "l1_sum_code", "l1_code_bwd", "l1_sum_code_pool", "l1_code_bwd_unpool", "l1_bwd", "l1_relu_bwd", "conv3x3_c3_fwd", "im2col3x3_c3",
    "col2im3x3_c3", "gen_tail_fwd", "gen_tail_compose", "gen_tail_bwd_act", "gen_tail_bwd", "copy_plan", "copy_multi",
]


def install(monkeypatch):
    """Replace the wrappers of b200lp.kernels with the emulations above and let the plugins run on the CPU."""
    import sys
    from b200lp import kernels as K
    from b200lp import lib as L
    this = sys.modules[__name__]
    for name in EMULATED:
        assert hasattr(K, name), name
        monkeypatch.setattr(K, name, getattr(this, name))
    monkeypatch.setattr(L, "require_device", lambda: 100)
