"""Tensor-level wrappers over the C ABI: allocate outputs / workspaces with torch, pass raw pointers.

No arithmetic happens in Python here.  Activations are NHWC float32 CUDA tensors of shape (N, H, W, C).
"""
from ctypes import byref, c_float, c_void_p

import torch

from . import lib as L


# bench.py's roofline pass: when PROFILE is a list, every C-ABI call is bracketed by CUDA events on the launching
# stream and recorded as (kernel family, algorithmic work {flops|bytes}, start event, end event).
PROFILE = None
# when WORK is a dict, every C-ABI call adds its algorithmic work to WORK[family] = {"flops", "bytes", "calls"} (host-side
# bookkeeping only; runners.holycow.GraphedTrainStep sets it while it captures the step, bench.py divides the replay's
# CUPTI kernel times by it)
WORK = None


class _timed:
    def __init__(self, family, flops=0.0, nbytes=0.0):
        self.family, self.work = family, {"flops": float(flops), "bytes": float(nbytes)}

    def __enter__(self):
        if WORK is not None:
            d = WORK.setdefault(self.family, {"flops": 0.0, "bytes": 0.0, "calls": 0})
            d["flops"] += self.work["flops"]
            d["bytes"] += self.work["bytes"]
            d["calls"] += 1
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.family, self.work, self.e0, e1))
        return False


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes) // 4, 1), dtype=torch.float32, device=device)


TF32, BF16X3 = 0, 1   # operand precisions of the tensor-core convolution (include/b200lp.h: `precision`)


def pack_conv_weight(w_oihw, scale=None, transpose=False, precision=TF32, out=None):
    """OIHW fp32 -> packed [Cout][tap][Cin] (or [Cin][tap'][Cout] when transpose), times *scale (device scalar):
    float32 holding tf32 values (precision TF32) or bfloat16 (2, ...) = (hi, lo) planes (precision BF16X3).
    `out`: re-pack into an existing buffer of the right shape."""
    lib = L.load()
    co, ci, kh, kw = w_oihw.shape
    assert kh == kw and kh in (1, 3)
    shape = (ci, kh * kw, co) if transpose else (co, kh * kw, ci)
    if precision != TF32:
        shape = (2,) + shape
    if out is None:
        out = torch.empty(shape, dtype=torch.float32 if precision == TF32 else torch.bfloat16, device=w_oihw.device)
    assert tuple(out.shape) == shape and out.is_contiguous(), (out.shape, shape)
    with _timed("pack_conv_weight", nbytes=8.0 * w_oihw.numel()):
        L.check(lib.b200lp_pack_conv_weight(L.ptr(w_oihw.contiguous()), L.ptr(scale), c_void_p(out.data_ptr()), co, ci,
                                            kh, 1 if transpose else 0, precision, L.stream_ptr()), "pack_conv_weight")
    return out


def pack_gconv_weight(w, transpose=False, precision=TF32, out=None):
    """Grouped 3x3 weight (C, cpg, 3, 3) -> block-diagonal dense tiles for conv_fwd(grouped=cpg): float32 (C, 9, 32)
    [TF32] or bfloat16 (2, C, 9, 64) [BF16X3]; transpose = data-gradient layout (flipped taps, roles swapped)."""
    lib = L.load()
    c, cpg, kh, kw = w.shape
    assert kh == 3 and kw == 3
    shape = (c, 9, 32) if precision == TF32 else (2, c, 9, 64)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32 if precision == TF32 else torch.bfloat16, device=w.device)
    assert tuple(out.shape) == shape and out.is_contiguous(), (out.shape, shape)
    with _timed("pack_conv_weight", nbytes=4.0 * w.numel() + out.numel() * out.element_size()):
        L.check(lib.b200lp_pack_gconv_weight(L.ptr(w.contiguous()), c_void_p(out.data_ptr()), c, cpg,
                                             1 if transpose else 0, precision, L.stream_ptr()), "pack_gconv_weight")
    return out


PACK_CHUNK = 16384


def pack_plan(rows, device):
    """Device-side table for `pack_conv_weight_multi`: rows = [(w_ptr, wp_ptr, Cout, Cin, taps, transpose, precision,
    numel), ...] (include/b200lp.h)."""
    tile_item, tile_index = [], []
    for i, r in enumerate(rows):
        assert r[4] <= 9, r
        tiles = ((r[2] + 31) // 32) * ((r[3] + 31) // 32)
        tile_item += [i] * tiles
        tile_index += list(range(tiles))
    return dict(table=torch.tensor(rows, dtype=torch.int64, device=device),
                tile_item=torch.tensor(tile_item, dtype=torch.int32, device=device),
                tile_index=torch.tensor(tile_index, dtype=torch.int32, device=device), n_tiles=len(tile_item),
                nbytes=8.0 * sum(r[7] for r in rows))


def pack_conv_weight_multi(plan):
    """One launch re-packs every copy of the plan: a block per 32 x 32 (co, ci) tile, coalesced on both sides."""
    lib = L.load()
    with _timed("pack_conv_weight", nbytes=plan["nbytes"]):
        L.check(lib.b200lp_pack_conv_weight_tiles(c_void_p(plan["table"].data_ptr()),
                                                  c_void_p(plan["tile_item"].data_ptr()),
                                                  c_void_p(plan["tile_index"].data_ptr()), plan["n_tiles"],
                                                  L.stream_ptr()), "pack_conv_weight_tiles")


def conv_fwd(x, wp, ksize, bias=None, residual=None, residual_mode=0, relu=False, round_tf32=False, block_n=0,
             out=None, emit_split=False, stages=0, scale=None, ctas_per_sm=0, splits=0, variant=0, a_stages=0,
             grouped=0):
    """x NHWC (N,H,W,Cin) float32 [TF32] or (2,N,H,W,Cin) bfloat16 (hi, lo) [BF16X3]; wp packed to match.
    `scale`: optional 1-element device tensor s, y = s * conv(x, wp) (+ bias ...).
    `grouped` = channels per group (0 = dense): wp from pack_gconv_weight, (Cout, 9, 32) / (2, Cout, 9, 64).
    Returns y (N,H,W,Cout) float32, or (y, y_split) with y_split (2,N,H,W,Cout) bfloat16 when emit_split."""
    lib = L.load()
    split_in = x.dtype == torch.bfloat16
    if split_in:
        assert x.dim() == 5 and wp.dtype == torch.bfloat16 and wp.dim() == 4, (x.shape, wp.shape)
        _, n, h, w, cin = x.shape
        cout = wp.shape[1]
        assert wp.shape[2] == ksize * ksize and wp.shape[3] == (64 if grouped else cin), (wp.shape, ksize, cin)
    else:
        n, h, w, cin = x.shape
        cout = wp.shape[0]
        assert wp.dtype == torch.float32 and wp.shape[1] == ksize * ksize and wp.shape[2] == (32 if grouped else cin), \
            (wp.shape, ksize, cin)
    kcin = grouped if grouped else cin          # algorithmic contraction length per tap
    y = out if out is not None else torch.empty((n, h, w, cout), dtype=torch.float32, device=x.device)
    y_split = torch.empty((2, n, h, w, cout), dtype=torch.bfloat16, device=x.device) if emit_split else None
    a = L.ConvArgs()
    a.x = L.ptr(x, x.dtype); a.wp = L.ptr(wp, wp.dtype); a.out_scale = L.ptr(scale); a.bias = L.ptr(bias)
    a.residual = L.ptr(residual)
    a.y = L.ptr(y); a.y_split = L.ptr(y_split, torch.bfloat16)
    a.N, a.H, a.W, a.Cin, a.Cout = n, h, w, cin, cout
    a.ksize = ksize
    a.residual_mode = residual_mode if residual is not None else 0
    a.relu = int(relu)
    a.round_tf32 = int(round_tf32)
    a.block_n = block_n
    a.stages = stages
    a.ctas_per_sm = ctas_per_sm
    a.precision = BF16X3 if split_in else TF32
    a.splits = splits
    a.variant, a.a_stages = variant, a_stages
    a.grouped = 1 if grouped else 0
    need = lib.b200lp_conv_fwd_workspace(byref(a))
    if need > 0:                      # few-tile layer: split-K partial sums
        ws = _ws(need, x.device)
        a.workspace, a.workspace_bytes = L.ptr(ws), ws.numel() * 4
    with _timed("resnext_grouped" if grouped else ("conv_igemm_bf16x3" if split_in else "conv_igemm_tf32"),
                flops=2.0 * n * h * w * kcin * cout * ksize * ksize):
        L.check(lib.b200lp_conv_fwd(byref(a), L.stream_ptr()), "conv_fwd")
    return (y, y_split) if emit_split else y


def wgrad_batch_pad(x, dy):
    """The weight-gradient kernel's K step is 32 pixels: on planes below 32 pixels one step spans 32 / (H*W) whole
    samples, so the batch must be a multiple of that (csrc/conv_wgrad.cu, plan_wgrad: pn).  A ragged batch (batch 1 on the
    generator's 4x4 planes, an odd batch on the discriminator's last blocks) is padded with all-zero samples: zero
    gradient rows contribute nothing to the sum over pixels.  Returns (x, dy) unchanged when no padding is needed."""
    n, h, w, _ = x.shape
    if h * w >= 32:
        return x, dy
    m = 32 // (h * w)
    pad = (-n) % m
    if pad == 0:
        return x, dy
    return (torch.cat((x, x.new_zeros((pad,) + tuple(x.shape[1:]))), 0),
            torch.cat((dy, dy.new_zeros((pad,) + tuple(dy.shape[1:]))), 0))


def conv_wgrad(x, dy, ksize, scale=1.0, kstep=0, stages=0, splits=0):
    """x (N,H,W,Cin), dy (N,H,W,Cout) -> dw OIHW (Cout,Cin,k,k) * scale."""
    lib = L.load()
    n_alg = x.shape[0]
    x, dy = wgrad_batch_pad(x, dy)
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    nbytes = lib.b200lp_conv_wgrad_workspace(n, h, w, cin, cout, ksize)
    if nbytes < 0:
        raise L.B200lpError(f"conv_wgrad_workspace: {L.last_error()}")
    if splits:
        nbytes = max(nbytes, splits * ksize * ksize * cin * cout * 4)
    ws = _ws(nbytes, x.device)
    dw = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    a = L.WgradArgs()
    a.x = L.ptr(x); a.dy = L.ptr(dy); a.dw = L.ptr(dw); a.workspace = L.ptr(ws)
    a.workspace_bytes = ws.numel() * 4
    a.N, a.H, a.W, a.Cin, a.Cout = n, h, w, cin, cout
    a.ksize = ksize
    a.scale = float(scale)
    a.kstep, a.stages, a.splits = kstep, stages, splits
    with _timed("conv_wgrad_tf32", flops=2.0 * n_alg * h * w * cin * cout * ksize * ksize):
        L.check(lib.b200lp_conv_wgrad(byref(a), L.stream_ptr()), "conv_wgrad")
    return dw


def conv_wgrad_sn_acc(x, dy, ksize, grad, weight=None, inv_sigma=None, u=None, v=None, accumulate=True):
    """grad (+)= s*G - s^2 <G, W> u v^T with G = wgrad(x, dy) (s = inv_sigma[0]; without inv_sigma: grad (+)= G).
    `grad`: contiguous OIHW fp32 buffer (the parameter's .grad view inside the gradient bucket) — written in place."""
    lib = L.load()
    n_alg = x.shape[0]
    x, dy = wgrad_batch_pad(x, dy)
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    assert tuple(grad.shape) == (cout, cin, ksize, ksize), (grad.shape, cout, cin, ksize)
    nbytes = lib.b200lp_conv_wgrad_sn_acc_workspace(n, h, w, cin, cout, ksize)
    if nbytes < 0:
        raise L.B200lpError(f"conv_wgrad_sn_acc_workspace: {L.last_error()}")
    ws = _ws(nbytes, x.device)
    a = L.WgradArgs()
    a.x = L.ptr(x); a.dy = L.ptr(dy); a.dw = L.ptr(grad); a.workspace = L.ptr(ws)
    a.workspace_bytes = ws.numel() * 4
    a.N, a.H, a.W, a.Cin, a.Cout = n, h, w, cin, cout
    a.ksize = ksize
    a.scale = 1.0
    wq = weight.detach() if weight is not None else None
    with _timed("conv_wgrad_tf32", flops=2.0 * n_alg * h * w * cin * cout * ksize * ksize):
        L.check(lib.b200lp_conv_wgrad_sn_acc(byref(a), L.ptr(wq), L.ptr(inv_sigma), L.ptr(u), L.ptr(v), int(accumulate),
                                             L.stream_ptr()), "conv_wgrad_sn_acc")
    return grad


EMA_CHUNK = 16384


def ema_plan(pairs):
    """Device table for ema_multi: pairs = [(avg, cur), ...] of equally shaped contiguous fp32 tensors on one device."""
    rows, chunk_t, chunk_o = [], [], []
    for i, (a_, c_) in enumerate(pairs):
        assert a_.is_contiguous() and c_.is_contiguous() and a_.shape == c_.shape and a_.dtype == c_.dtype == torch.float32
        n = a_.numel()
        rows.append([c_.data_ptr(), 0, 0, 0, a_.data_ptr(), n])          # OptTensor {p, g, m, v, ema, n}
        for o in range(0, n, EMA_CHUNK):
            chunk_t.append(i)
            chunk_o.append(o)
    dev = pairs[0][0].device
    return dict(table=torch.tensor(rows, dtype=torch.int64, device=dev),
                chunk_t=torch.tensor(chunk_t, dtype=torch.int32, device=dev),
                chunk_o=torch.tensor(chunk_o, dtype=torch.int64, device=dev), n_chunks=len(chunk_t),
                sig=tuple((r[0], r[4], r[5]) for r in rows), nbytes=12.0 * sum(r[5] for r in rows))


def ema_multi(plan, alpha):
    """avg = avg * alpha + cur * (1 - alpha) for every pair of the plan, one launch (runners/holycow.py:99-105)."""
    lib = L.load()
    with _timed("optimizer", nbytes=plan["nbytes"]):
        L.check(lib.b200lp_ema_multi(c_void_p(plan["table"].data_ptr()), c_void_p(plan["chunk_t"].data_ptr()),
                                     c_void_p(plan["chunk_o"].data_ptr()), plan["n_chunks"], EMA_CHUNK, c_float(alpha),
                                     L.stream_ptr()), "ema_multi")


def copy_plan(pairs):
    """Device table for copy_multi: pairs = [(dst, src), ...] of equally sized contiguous tensors on one device."""
    rows = []
    for d, s_ in pairs:
        assert d.is_contiguous() and s_.is_contiguous() and d.numel() * d.element_size() == s_.numel() * s_.element_size()
        if d.numel():
            rows.append((d.data_ptr(), s_.data_ptr(), d.numel() * d.element_size()))
    dev = pairs[0][0].device
    return dict(table=torch.tensor(rows, dtype=torch.int64, device=dev), count=len(rows), sig=tuple(rows),
                nbytes=2.0 * sum(r[2] for r in rows))


def copy_multi(plan):
    lib = L.load()
    if plan["count"] == 0:
        return
    with _timed("elementwise", nbytes=plan["nbytes"]):
        L.check(lib.b200lp_copy_multi(c_void_p(plan["table"].data_ptr()), plan["count"], L.stream_ptr()), "copy_multi")


def in_stats(x, eps):
    """x (N,H,W,C) -> mean (N,C), rstd (N,C) of each (n,c) plane (biased variance)."""
    lib = L.load()
    n, h, w, c = x.shape
    ws = _ws(lib.b200lp_in_stats_workspace(n, h * w, c), x.device)
    mean = torch.empty((n, c), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    with _timed("in_stats", nbytes=4.0 * x.numel()):
        L.check(lib.b200lp_in_stats(L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(ws), ws.numel() * 4, n, h * w, c,
                                    c_float(eps), L.stream_ptr()), "in_stats")
    return mean, rstd


def _affine_views(gamma, beta):
    # gamma / beta are (N, C) views into the projector output (row stride = affine_stride, unit column stride)
    assert gamma.stride(1) == 1 and beta.stride(1) == 1 and gamma.stride(0) == beta.stride(0)
    assert gamma.dtype == torch.float32 and gamma.is_cuda
    return c_void_p(gamma.data_ptr()), c_void_p(beta.data_ptr()), gamma.stride(0)


def adain_relu(x, mean, rstd, gamma, beta, upsample2=False, round_tf32=True, want_f32=True, want_split=False):
    """Returns y (fp32) / (y, y_split) / y_split according to want_f32 / want_split; y_split is (2,N,H',W',C) bf16."""
    lib = L.load()
    n, h, w, c = x.shape
    gp, bp, stride = _affine_views(gamma, beta)
    s = 2 if upsample2 else 1
    y = torch.empty((n, h * s, w * s, c), dtype=torch.float32, device=x.device) if want_f32 else None
    ys = torch.empty((2, n, h * s, w * s, c), dtype=torch.bfloat16, device=x.device) if want_split else None
    out_elems = n * h * s * w * s * c
    with _timed("adain_relu", nbytes=4.0 * (x.numel() + out_elems * (int(want_f32) + int(want_split)))):
        L.check(lib.b200lp_adain_relu(L.ptr(x), L.ptr(mean), L.ptr(rstd), gp, bp, stride, L.ptr(y),
                                      L.ptr(ys, torch.bfloat16), n, h, w, c, int(upsample2), int(round_tf32),
                                      L.stream_ptr()), "adain_relu")
    if want_f32 and want_split:
        return y, ys
    return y if want_f32 else ys


_SYNC_BUFFERS = {}


def _sync_buffer(device, n):
    """Zeroed uint32 counters of the per-sample barriers (the kernels leave them zero); one buffer per (device, stream)."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    buf = _SYNC_BUFFERS.get(key)
    if buf is None or buf.numel() < 4 * n:
        buf = torch.zeros(max(4 * n, 512), dtype=torch.int32, device=device)
        _SYNC_BUFFERS[key] = buf
    return buf


def adain_stats_apply(x, gamma, beta, eps, upsample2=False, round_tf32=True, want_f32=True, want_split=False, fused=None):
    """Instance-norm statistics + AdaIN + ReLU (+2x) of one site -> (mean, rstd, outputs as adain_relu returns them).
    Default: in_stats (2 launches) + adain_relu.  `fused=True` / B200LP_ADAIN_FUSED=1: ONE launch
    (b200lp_adain_relu_fused: partials, per-sample barrier, channel-sliced merge, barrier, apply).  The single launch was
    measured SLOWER inside the step (17 sites: 1.03 ms vs 0.65 ms) — two grid-wide waits per site cost
    more than the launch gaps they replace — so it is kept as a tested, opt-in form only."""
    import os
    lib = L.load()
    n, h, w, c = x.shape
    if fused is None:
        fused = bool(os.environ.get("B200LP_ADAIN_FUSED"))
    if not fused:
        mean, rstd = in_stats(x, eps)
        return mean, rstd, adain_relu(x, mean, rstd, gamma, beta, upsample2=upsample2, round_tf32=round_tf32,
                                      want_f32=want_f32, want_split=want_split)
    gp, bp, stride = _affine_views(gamma, beta)
    sc = 2 if upsample2 else 1
    ws = _ws(lib.b200lp_in_stats_workspace(n, h * w, c), x.device)
    mean = torch.empty((n, c), dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    y = torch.empty((n, h * sc, w * sc, c), dtype=torch.float32, device=x.device) if want_f32 else None
    ys = torch.empty((2, n, h * sc, w * sc, c), dtype=torch.bfloat16, device=x.device) if want_split else None
    sync = _sync_buffer(x.device, n)
    out_elems = n * h * sc * w * sc * c
    with _timed("adain_relu", nbytes=4.0 * (x.numel() + out_elems * (int(want_f32) + int(want_split)))):
        rc = lib.b200lp_adain_relu_fused(L.ptr(x), gp, bp, stride, L.ptr(y), L.ptr(ys, torch.bfloat16), L.ptr(mean),
                                         L.ptr(rstd), L.ptr(ws), ws.numel() * 4, L.ptr(sync, torch.int32), n, h, w, c,
                                         c_float(eps), int(upsample2), int(round_tf32), L.stream_ptr())
    if rc != 0:      # grid not co-resident on this device: the two-kernel form (nothing was launched)
        mean, rstd = in_stats(x, eps)
        return mean, rstd, adain_relu(x, mean, rstd, gamma, beta, upsample2=upsample2, round_tf32=round_tf32,
                                      want_f32=want_f32, want_split=want_split)
    out = (y, ys) if (want_f32 and want_split) else (y if want_f32 else ys)
    return mean, rstd, out


def adain_relu_bwd(x, mean, rstd, gamma, beta, dy, upsample2=False, add=None, round_tf32=False):
    """-> (dx, dgamma, dbeta); `add`: a second gradient of x merged into dx, `round_tf32`: dx stored rounded to tf32."""
    lib = L.load()
    n, h, w, c = x.shape
    gp, bp, stride = _affine_views(gamma, beta)
    ws = _ws(lib.b200lp_adain_relu_bwd_workspace(n, h * w, c), x.device)
    dx = torch.empty_like(x)
    dgamma = torch.empty((n, c), dtype=torch.float32, device=x.device)
    dbeta = torch.empty_like(dgamma)
    with _timed("adain_relu_bwd", nbytes=4.0 * (2 * x.numel() + dy.numel())):   # ideal: read x, dy once; write dx
        L.check(lib.b200lp_adain_relu_bwd(L.ptr(x), L.ptr(mean), L.ptr(rstd), gp, bp, stride, L.ptr(dy), L.ptr(dx),
                                          L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), ws.numel() * 4, n, h, w, c,
                                          int(upsample2), L.ptr(add), int(round_tf32), L.stream_ptr()), "adain_relu_bwd")
    return dx, dgamma, dbeta


def nchw_to_nhwc(x):
    lib = L.load()
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
    L.check(lib.b200lp_nchw_to_nhwc(L.ptr(x), L.ptr(y), n, c, h, w, L.stream_ptr()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x):
    lib = L.load()
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    L.check(lib.b200lp_nhwc_to_nchw(L.ptr(x), L.ptr(y), n, c, h, w, L.stream_ptr()), "nhwc_to_nchw")
    return y


def relu_round(x):
    lib = L.load()
    y = torch.empty_like(x)
    with _timed("elementwise", nbytes=8.0 * x.numel()):
        L.check(lib.b200lp_relu_round(L.ptr(x), L.ptr(y), x.numel(), L.stream_ptr()), "relu_round")
    return y


def relu_bwd(y, dy):
    lib = L.load()
    dx = torch.empty_like(dy)
    with _timed("elementwise", nbytes=12.0 * dy.numel()):
        L.check(lib.b200lp_relu_bwd(L.ptr(y), L.ptr(dy), L.ptr(dx), dy.numel(), L.stream_ptr()), "relu_bwd")
    return dx


def relu_bwd_fused(y, dy, add=None, want_quarter=False, bias_a=None, bias_b=None, round_tf32=False):
    """dx = [y > 0] * (dy + add); optionally dq = 0.25 * dx; bias_a / bias_b (C,) += column sums of dx (in place);
    round_tf32: dx / dq stored rounded to tf32 (operands of the following data / weight gradient MMAs).
    Returns dx or (dx, dq)."""
    lib = L.load()
    c = y.shape[-1]
    dx = torch.empty_like(y)
    dq = torch.empty_like(y) if want_quarter else None
    nb = 12.0 + (4.0 if add is not None else 0.0) + (4.0 if want_quarter else 0.0)
    with _timed("elementwise", nbytes=nb * y.numel()):
        L.check(lib.b200lp_relu_bwd_fused(L.ptr(y), L.ptr(dy), L.ptr(add), L.ptr(dx), L.ptr(dq), L.ptr(bias_a),
                                          L.ptr(bias_b), y.numel() // c, c, int(round_tf32), L.stream_ptr()),
                "relu_bwd_fused")
    return (dx, dq) if want_quarter else dx


def avgpool2(x, addend=None, round_tf32=False):
    lib = L.load()
    n, h2, w2, c = x.shape
    y = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.float32, device=x.device)
    with _timed("elementwise", nbytes=4.0 * (x.numel() + y.numel())):
        L.check(lib.b200lp_avgpool2(L.ptr(x), L.ptr(addend), L.ptr(y), n, h2 // 2, w2 // 2, c, int(round_tf32),
                                    L.stream_ptr()), "avgpool2")
    return y


def avgpool2_bwd(dy):
    lib = L.load()
    n, h, w, c = dy.shape
    dx = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=dy.device)
    with _timed("elementwise", nbytes=4.0 * (dy.numel() + dx.numel())):
        L.check(lib.b200lp_avgpool2_bwd(L.ptr(dy), L.ptr(dx), n, h, w, c, L.stream_ptr()), "avgpool2_bwd")
    return dx


def upsample2_bwd(dy):
    lib = L.load()
    n, h2, w2, c = dy.shape
    dx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.float32, device=dy.device)
    with _timed("elementwise", nbytes=4.0 * (dy.numel() + dx.numel())):
        L.check(lib.b200lp_upsample2_bwd(L.ptr(dy), L.ptr(dx), n, h2 // 2, w2 // 2, c, L.stream_ptr()), "upsample2_bwd")
    return dx


def l1_sum(a, b, out, scale):
    """out[0] += scale * sum|a-b| (out: 1-element device tensor, caller zeroes it)."""
    lib = L.load()
    with _timed("l1", nbytes=8.0 * a.numel()):
        L.check(lib.b200lp_l1_sum(L.ptr(a), L.ptr(b), L.ptr(out), a.numel(), c_float(scale), L.stream_ptr()), "l1_sum")


def l1_sum_code(a, b, out, scale):
    """out[0] += scale * sum|a-b|; returns the backward code tensor (uint8, a.numel() / 4 bytes: 2 bits per element, ReLU mask
    of `a` and sign(a-b)) that `l1_code_bwd` consumes instead of the two feature maps."""
    lib = L.load()
    code = torch.empty((a.numel() // 4,), dtype=torch.uint8, device=a.device)
    with _timed("l1", nbytes=8.25 * a.numel()):
        L.check(lib.b200lp_l1_sum_code(L.ptr(a), L.ptr(b), L.ptr(out), L.ptr(code, torch.uint8), a.numel(), c_float(scale),
                                       L.stream_ptr()), "l1_sum_code")
    return code


def l1_code_bwd(code, shape, gscale, scale2, d_in=None):
    """d_out = mask * tf32(d_in + sign * gscale[0] * scale2) from the code of l1_sum_code; `shape`: the feature's shape."""
    lib = L.load()
    d = torch.empty(shape, dtype=torch.float32, device=code.device)
    assert d.numel() == 4 * code.numel()
    with _timed("l1", nbytes=(4.25 + (4.0 if d_in is not None else 0.0)) * d.numel()):
        L.check(lib.b200lp_l1_code_bwd(L.ptr(code, torch.uint8), L.ptr(gscale), c_float(scale2), L.ptr(d_in), L.ptr(d),
                                       d.numel(), L.stream_ptr()), "l1_code_bwd")
    return d


def l1_sum_code_pool(a, b, out, scale, ap=None, bp=None):
    """l1_sum_code fused with the 2x2 average pool that follows the tap: -> (code, a_pooled, b_pooled), pooled maps tf32-rounded
    (written into `ap` / `bp` when given: contiguous (N, H/2, W/2, C) buffers, e.g. the two halves of one batch tensor)."""
    lib = L.load()
    n, h, w, c = a.shape
    code = torch.empty((a.numel() // 4,), dtype=torch.uint8, device=a.device)
    if ap is None:
        ap = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=a.device)
    if bp is None:
        bp = torch.empty_like(ap)
    assert tuple(ap.shape) == tuple(bp.shape) == (n, h // 2, w // 2, c)
    with _timed("l1", nbytes=8.25 * a.numel() + 8.0 * ap.numel()):
        L.check(lib.b200lp_l1_sum_code_pool(L.ptr(a), L.ptr(b), L.ptr(out), L.ptr(code, torch.uint8), L.ptr(ap), L.ptr(bp),
                                            n, h, w, c, c_float(scale), L.stream_ptr()), "l1_sum_code_pool")
    return code, ap, bp


def l1_code_bwd_unpool(code, shape, gscale, scale2, d_low):
    """avgpool2_bwd + l1_code_bwd in one pass: d_low is the gradient of the pooled map, `shape` the tap's (N,H,W,C)."""
    lib = L.load()
    n, h, w, c = shape
    d = torch.empty(shape, dtype=torch.float32, device=code.device)
    assert tuple(d_low.shape) == (n, h // 2, w // 2, c)
    with _timed("l1", nbytes=4.25 * d.numel() + 4.0 * d_low.numel()):
        L.check(lib.b200lp_l1_code_bwd_unpool(L.ptr(code, torch.uint8), L.ptr(gscale), c_float(scale2), L.ptr(d_low),
                                              L.ptr(d), n, h, w, c, L.stream_ptr()), "l1_code_bwd_unpool")
    return d


def l1_bwd(a, b, gscale, scale2, da=None):
    """da (+)= sign(a-b) * gscale[0] * scale2."""
    lib = L.load()
    acc = da is not None
    if da is None:
        da = torch.empty_like(a)
    with _timed("l1", nbytes=12.0 * a.numel()):
        L.check(lib.b200lp_l1_bwd(L.ptr(a), L.ptr(b), L.ptr(gscale), c_float(scale2), L.ptr(da), a.numel(), int(acc),
                                  L.stream_ptr()), "l1_bwd")
    return da


def l1_relu_bwd(a, b, gscale, scale2, d_in=None):
    """d_out = [a > 0] * (d_in + sign(a-b) * gscale[0]*scale2): one VGG backward tap (L1 term + ReLU mask) in one pass."""
    lib = L.load()
    d_out = torch.empty_like(a)
    with _timed("l1", nbytes=(16.0 if d_in is not None else 12.0) * a.numel()):
        L.check(lib.b200lp_l1_relu_bwd(L.ptr(a), L.ptr(b), L.ptr(gscale), c_float(scale2), L.ptr(d_in), L.ptr(d_out),
                                       a.numel(), L.stream_ptr()), "l1_relu_bwd")
    return d_out


def conv3x3_c3_fwd(x_nchw, w, wscale=None, bias=None, pre_scale=None, pre_shift=None, relu=False, round_tf32=False,
                   tensor_cores=None):
    """3x3 conv of a 3-channel NCHW image -> NHWC features.  `tensor_cores` (default: Cout == 64, unless
    B200LP_C3_CUDA_CORES=1): the tcgen05 implicit GEMM with tf32 operands instead of the FP32 CUDA-core kernel."""
    import os
    lib = L.load()
    n, c, h, wd = x_nchw.shape
    assert c == 3
    cout = w.shape[0]
    if tensor_cores is None:
        tensor_cores = cout == 64 and not os.environ.get("B200LP_C3_CUDA_CORES")
    y = torch.empty((n, h, wd, cout), dtype=torch.float32, device=x_nchw.device)
    fn = lib.b200lp_conv3x3_c3_fwd_tc if tensor_cores else lib.b200lp_conv3x3_c3_fwd
    with _timed("conv3x3_c3_fwd", nbytes=4.0 * (x_nchw.numel() + y.numel())):
        L.check(fn(L.ptr(x_nchw), L.ptr(w.contiguous()), L.ptr(wscale), L.ptr(bias), L.ptr(pre_scale), L.ptr(pre_shift),
                   L.ptr(y), n, h, wd, cout, int(relu), int(round_tf32), L.stream_ptr()), "conv3x3_c3_fwd")
    return y


def conv3x3_c3_dgrad(dy, w, wscale=None, pre_scale=None):
    lib = L.load()
    n, h, wd, cout = dy.shape
    dx = torch.empty((n, 3, h, wd), dtype=torch.float32, device=dy.device)
    with _timed("conv3x3_c3_dgrad", nbytes=4.0 * (dy.numel() + dx.numel())):
        L.check(lib.b200lp_conv3x3_c3_dgrad(L.ptr(dy), L.ptr(w), L.ptr(wscale), L.ptr(pre_scale), L.ptr(dx), n, h, wd,
                                            cout, L.stream_ptr()), "conv3x3_c3_dgrad")
    return dx


def conv3x3_c3_wgrad(x_nchw, dy, scale=1.0):
    lib = L.load()
    n, h, wd, cout = dy.shape
    dw = torch.empty((cout, 3, 3, 3), dtype=torch.float32, device=dy.device)
    with _timed("conv3x3_c3_wgrad", nbytes=4.0 * (dy.numel() + x_nchw.numel())):
        L.check(lib.b200lp_conv3x3_c3_wgrad(L.ptr(x_nchw), L.ptr(dy), L.ptr(dw), c_float(scale), n, h, wd, cout,
                                            L.stream_ptr()), "conv3x3_c3_wgrad")
    return dw


def gen_tail_fwd(x, w, wscale, bias):
    lib = L.load()
    n, h, wd, cin = x.shape
    rgbs = torch.empty((n, 3, h, wd), dtype=torch.float32, device=x.device)
    segm = torch.empty((n, 1, h, wd), dtype=torch.float32, device=x.device)
    t = torch.empty((n, h, wd, 4), dtype=torch.float32, device=x.device)
    with _timed("gen_tail_fwd", nbytes=4.0 * (x.numel() + rgbs.numel() + segm.numel() + t.numel())):
        L.check(lib.b200lp_gen_tail_fwd(L.ptr(x), L.ptr(w), L.ptr(wscale), L.ptr(bias), L.ptr(rgbs), L.ptr(segm),
                                        L.ptr(t), n, h, wd, cin, L.stream_ptr()), "gen_tail_fwd")
    return rgbs, segm, t


def gen_tail_compose(a32, bias):
    """a32 (N,H,W,32): tensor-core conv output whose first 4 channels are the tail's pre-tanh values (bias not yet added)
    -> fake_rgbs (N,3,H,W), fake_segm (N,1,H,W), t (N,H,W,4)."""
    lib = L.load()
    n, h, wd, stride = a32.shape
    rgbs = torch.empty((n, 3, h, wd), dtype=torch.float32, device=a32.device)
    segm = torch.empty((n, 1, h, wd), dtype=torch.float32, device=a32.device)
    t = torch.empty((n, h, wd, 4), dtype=torch.float32, device=a32.device)
    with _timed("gen_tail_fwd", nbytes=4.0 * (a32.numel() // 8 * 2 + rgbs.numel() + segm.numel() + t.numel())):
        L.check(lib.b200lp_gen_tail_compose(L.ptr(a32), L.ptr(bias), L.ptr(rgbs), L.ptr(segm), L.ptr(t), n, h, wd, stride,
                                            L.stream_ptr()), "gen_tail_compose")
    return rgbs, segm, t


def gen_tail_bwd_act(t, d_rgbs, d_segm, stride=32):
    """Pre-tanh gradient (N,H,W,stride): 4 real channels, zero padded (and tf32-rounded) when stride is 32."""
    lib = L.load()
    n, h, wd, _ = t.shape
    da = torch.empty((n, h, wd, stride), dtype=torch.float32, device=t.device)
    with _timed("gen_tail_bwd_act", nbytes=4.0 * (da.numel() + 2 * t.numel())):
        L.check(lib.b200lp_gen_tail_bwd_act(L.ptr(t), L.ptr(d_rgbs), L.ptr(d_segm), L.ptr(da), n, h, wd, stride,
                                            L.stream_ptr()), "gen_tail_bwd_act")
    return da


def gen_tail_bwd(x, t, w, wscale, d_rgbs, d_segm, need_dx=True, need_dw=True):
    """Backward of the generator tail.  Returns (dx, g, db): g = gradient w.r.t. the scaled weight (w*wscale), OIHW.
    The weight gradient runs on the tensor cores: the 4-channel pre-tanh gradient is written as a zero-padded
    32-channel NHWC tensor and handed to conv_wgrad."""
    lib = L.load()
    n, h, wd, cin = x.shape
    stride = 32 if need_dw else 4
    da = torch.empty((n, h, wd, stride), dtype=torch.float32, device=x.device)
    with _timed("gen_tail_bwd_act", nbytes=4.0 * (da.numel() + 2 * t.numel())):
        L.check(lib.b200lp_gen_tail_bwd_act(L.ptr(t), L.ptr(d_rgbs), L.ptr(d_segm), L.ptr(da), n, h, wd, stride,
                                            L.stream_ptr()), "gen_tail_bwd_act")
    dx = dw = db = None
    if need_dx:
        dx = torch.empty_like(x)
        with _timed("gen_tail_bwd_data", nbytes=4.0 * (da.numel() + dx.numel())):
            L.check(lib.b200lp_gen_tail_bwd_data(L.ptr(da), L.ptr(w), L.ptr(wscale), L.ptr(dx), n, h, wd, cin, stride,
                                                 L.stream_ptr()), "gen_tail_bwd_data")
    if need_dw:
        dw = conv_wgrad(x, da, 3)[:4].contiguous()
        db = bias_grad(da)[:4].contiguous()
    return dx, dw, db


def im2col3x3_c3(x_nchw):
    """(N,3,H,W) image -> (N,H,W,32) patch matrix (27 columns c*9+kh*3+kw, 5 zero columns), tf32-rounded."""
    lib = L.load()
    n, c, h, w = x_nchw.shape
    assert c == 3
    col = torch.empty((n, h, w, 32), dtype=torch.float32, device=x_nchw.device)
    with _timed("c3_im2col", nbytes=4.0 * (x_nchw.numel() + col.numel())):
        L.check(lib.b200lp_im2col3x3_c3(L.ptr(x_nchw), L.ptr(col), n, h, w, L.stream_ptr()), "im2col3x3_c3")
    return col


def col2im3x3_c3(dcol, pre_scale=None):
    lib = L.load()
    n, h, w, c = dcol.shape
    assert c == 32
    dx = torch.empty((n, 3, h, w), dtype=torch.float32, device=dcol.device)
    with _timed("c3_col2im", nbytes=4.0 * (dcol.numel() + dx.numel())):
        L.check(lib.b200lp_col2im3x3_c3(L.ptr(dcol), L.ptr(pre_scale), L.ptr(dx), n, h, w, L.stream_ptr()),
                "col2im3x3_c3")
    return dx


def conv3x3_c3_dgrad_tc(dy, w_t_packed, wscale=None, pre_scale=None):
    """Data gradient of a Cin=3 conv on the tensor cores: dcol = conv1x1(dy; W^T) (Cout -> 32), dx = col2im(dcol).
    `w_t_packed` = pack_conv_weight of the (32, Cout, 1, 1) matrix made by c3_transposed_weight()."""
    dcol = conv_fwd(dy, w_t_packed, 1, scale=wscale)
    return col2im3x3_c3(dcol, pre_scale)


def c3_transposed_weight(w_oihw):
    """(Cout,3,3,3) -> (32, Cout, 1, 1): row t = c*9+kh*3+kw holds w[:, c, kh, kw]; rows 27..31 are zero."""
    cout = w_oihw.shape[0]
    wt = torch.zeros((32, cout), dtype=torch.float32, device=w_oihw.device)
    wt[:27] = w_oihw.detach().reshape(cout, 27).t()
    return pack_conv_weight(wt.reshape(32, cout, 1, 1))


def conv3x3_c3_wgrad_tc(x_nchw, dy):
    """Weight gradient of a Cin=3 conv on the tensor cores: conv_wgrad(im2col(x), dy, 1x1)[:, :27]."""
    cout = dy.shape[-1]
    g = conv_wgrad(im2col3x3_c3(x_nchw), dy, 1)          # (Cout, 32, 1, 1)
    return g.reshape(cout, 32)[:, :27].reshape(cout, 3, 3, 3).contiguous()


def bias_grad(dy, acc_into=None):
    """db[c] = sum over pixels of dy[..., c]; with `acc_into` (the bias parameter's gradient buffer) db is added there."""
    lib = L.load()
    c = dy.shape[-1]
    if acc_into is not None:
        assert acc_into.numel() == c
        with _timed("elementwise", nbytes=4.0 * dy.numel()):
            L.check(lib.b200lp_bias_grad_acc(L.ptr(dy), L.ptr(acc_into), dy.numel() // c, c, L.stream_ptr()),
                    "bias_grad_acc")
        return acc_into
    db = torch.empty((c,), dtype=torch.float32, device=dy.device)
    with _timed("elementwise", nbytes=4.0 * dy.numel()):
        L.check(lib.b200lp_bias_grad(L.ptr(dy), L.ptr(db), dy.numel() // c, c, L.stream_ptr()), "bias_grad")
    return db


def sn_sigma_multi(layers, training):
    """Batched spectral-norm pass over `layers` = list of (weight_orig, weight_u, weight_v, eps, scratch) with
    weight viewed as [rows][cols].  Returns (inv_sigma (T,), [(snap_u, snap_v)]) — one launch triple for all of them."""
    lib = L.load()
    dev = layers[0][0].device
    total = sum(w.shape[0] + w[0].numel() for w, *_ in layers)
    snap = torch.empty(total, dtype=torch.float32, device=dev)
    inv = torch.empty(len(layers), dtype=torch.float32, device=dev)
    items = (L.SnItem * len(layers))()
    snaps = []
    off = 0
    for i, (w, u, v, eps, scratch) in enumerate(layers):
        rows, cols = w.shape[0], w[0].numel()
        su, sv = snap[off:off + rows], snap[off + rows:off + rows + cols]
        off += rows + cols
        it = items[i]
        it.w, it.u, it.v = w.data_ptr(), u.data_ptr(), v.data_ptr()
        it.snap_u, it.snap_v = su.data_ptr(), sv.data_ptr()
        it.scratch = scratch.data_ptr()
        it.inv_sigma = inv.data_ptr() + 4 * i
        it.rows, it.cols, it.eps = rows, cols, eps
        snaps.append((su, sv))
    with _timed("spectral_norm", nbytes=8.0 * sum(w.numel() for w, *_ in layers)):
        L.check(lib.b200lp_sn_sigma_multi(items, len(layers), int(training), L.stream_ptr()), "sn_sigma_multi")
    return inv, snaps


def sn_scratch(w):
    lib = L.load()
    n = lib.b200lp_sn_scratch_floats(w.shape[0], w[0].numel())
    return torch.empty(n, dtype=torch.float32, device=w.device)


def sn_wgrad_fix(g, w, inv_sigma, u, v):
    """dw = s*g - s^2 <g,w> u v^T (all tensors shaped like w)."""
    lib = L.load()
    rows, cols = w.shape[0], w[0].numel()
    ws = _ws(lib.b200lp_sn_wgrad_fix_workspace(w.numel()), w.device)
    dw = torch.empty_like(w)
    with _timed("spectral_norm", nbytes=16.0 * w.numel()):
        L.check(lib.b200lp_sn_wgrad_fix(L.ptr(g), L.ptr(w.detach().contiguous()), L.ptr(inv_sigma), L.ptr(u), L.ptr(v),
                                        L.ptr(dw), L.ptr(ws), rows, cols, L.stream_ptr()), "sn_wgrad_fix")
    return dw


# ---------------------------------------------------------------------------------------------------------------------
# Pose encoder (MobileNetV2) forward kernels — csrc/mobilenet.cu
# ---------------------------------------------------------------------------------------------------------------------
def pw_conv(x2d, weight, in_scale=None, in_shift=None, in_relu6=False, bias=None, want_stats=False):
    """y (M, Cout) = f(x2d (M, Cin)) @ weight (Cout, Cin)^T (+ bias); f = producer BatchNorm (+ReLU6) applied on load.
    Returns y or (y, part) with part (parts, 2, Cout) per-channel sum / sum-of-squares partials of y."""
    lib = L.load()
    m, cin = x2d.shape
    cout = weight.shape[0]
    assert weight.numel() == cout * cin
    y = torch.empty((m, cout), dtype=torch.float32, device=x2d.device)
    part = None
    if want_stats:
        part = torch.empty((lib.b200lp_pw_conv_parts(m, cout), 2, cout), dtype=torch.float32, device=x2d.device)
    need = lib.b200lp_pw_conv_workspace(m, cin, cout)       # > 0: few-tile layer, split-K partial sums
    ws = _ws(need, x2d.device) if need > 0 else None
    with _timed("pose_encoder", flops=2.0 * m * cin * cout):
        L.check(lib.b200lp_pw_conv_ws(L.ptr(x2d), L.ptr(in_scale), L.ptr(in_shift), int(in_relu6), L.ptr(weight),
                                      L.ptr(bias), L.ptr(y), L.ptr(part), m, cin, cout, L.ptr(ws),
                                      ws.numel() * 4 if ws is not None else 0, L.stream_ptr()), "pw_conv")
    return (y, part) if want_stats else y


def dw_conv3x3(x, weight, in_scale, in_shift, stride, want_stats=False):
    """Depthwise 3x3 (padding 1) on relu6(x*scale+shift); x (N,H,W,C) raw producer output -> y (N,Ho,Wo,C) raw."""
    lib = L.load()
    n, h, w, c = x.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device)
    part = None
    if want_stats:
        part = torch.empty((lib.b200lp_dw_conv3x3_parts(n, h, w, stride), 2, c), dtype=torch.float32, device=x.device)
    with _timed("pose_encoder", nbytes=4.0 * (x.numel() + y.numel())):
        L.check(lib.b200lp_dw_conv3x3(L.ptr(x), L.ptr(in_scale), L.ptr(in_shift), L.ptr(weight), L.ptr(y), L.ptr(part),
                                      n, h, w, c, stride, L.stream_ptr()), "dw_conv3x3")
    return (y, part) if want_stats else y


def mbv2_stem(x_nchw, weight, want_stats=False):
    """3x3 stride-2 conv 3 -> 32 on the NCHW image -> (N, H/2, W/2, 32) NHWC raw."""
    lib = L.load()
    n, c, h, w = x_nchw.shape
    assert c == 3 and tuple(weight.shape) == (32, 3, 3, 3)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, ho, wo, 32), dtype=torch.float32, device=x_nchw.device)
    part = None
    if want_stats:
        part = torch.empty((lib.b200lp_mbv2_stem_parts(n, h, w), 2, 32), dtype=torch.float32, device=x_nchw.device)
    with _timed("pose_encoder", nbytes=4.0 * (x_nchw.numel() + y.numel())):
        L.check(lib.b200lp_mbv2_stem(L.ptr(x_nchw), L.ptr(weight), L.ptr(y), L.ptr(part), n, h, w, L.stream_ptr()),
                "mbv2_stem")
    return (y, part) if want_stats else y


def bn_finalize(bn, part, count, training, want_stats=False):
    """(scale, shift) of a torch BatchNorm2d module `bn` for the kernels' on-load normalisation.  training: batch
    statistics from `part` (+ running-statistics update in place, like the module's forward); else running stats.
    want_stats: also return (mean, rstd) the layer normalised with (inputs of bn_bwd)."""
    lib = L.load()
    c = bn.num_features
    dev = bn.weight.device
    scale = torch.empty(c, dtype=torch.float32, device=dev)
    shift = torch.empty(c, dtype=torch.float32, device=dev)
    mean = torch.empty(c, dtype=torch.float32, device=dev) if want_stats else None
    rstd = torch.empty(c, dtype=torch.float32, device=dev) if want_stats else None
    track = bn.track_running_stats and bn.running_mean is not None
    nbt = bn.num_batches_tracked if (training and track and bn.num_batches_tracked is not None) else None
    L.check(lib.b200lp_bn_finalize(L.ptr(part), part.shape[0] if part is not None else 0, int(count),
                                   L.ptr(bn.weight.detach()), L.ptr(bn.bias.detach()),
                                   L.ptr(bn.running_mean) if track else None, L.ptr(bn.running_var) if track else None,
                                   L.ptr(nbt, torch.int64), c_float(bn.momentum if bn.momentum is not None else 0.1),
                                   c_float(bn.eps), L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(rstd), c,
                                   int(training), L.stream_ptr()),
            "bn_finalize")
    return (scale, shift, mean, rstd) if want_stats else (scale, shift)


def bn_apply(x, scale, shift, residual=None, relu6=False):
    lib = L.load()
    c = x.shape[-1]
    y = torch.empty_like(x)
    with _timed("pose_encoder", nbytes=4.0 * x.numel() * (3 if residual is not None else 2)):
        L.check(lib.b200lp_bn_apply(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(residual), L.ptr(y), x.numel() // c, c,
                                    int(relu6), L.stream_ptr()), "bn_apply")
    return y


def bn_relu6_avgpool(x, scale, shift):
    """x (N,H,W,C) raw -> (N, C): spatial mean of relu6(x*scale+shift)."""
    lib = L.load()
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=torch.float32, device=x.device)
    with _timed("pose_encoder", nbytes=4.0 * x.numel()):
        L.check(lib.b200lp_bn_relu6_avgpool(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(y), n, h * w, c, L.stream_ptr()),
                "bn_relu6_avgpool")
    return y


# ---------------------------------------------------------------------------------------------------------------------
# Identity encoder (ResNeXt50-32x4d) + shared BatchNorm backward — csrc/encoder.cu
# ---------------------------------------------------------------------------------------------------------------------
def col_stats(x2d):
    """x (M, C) -> part (parts, 2, C): per-channel sum / sum-of-squares partials for bn_finalize."""
    lib = L.load()
    m, c = x2d.shape
    part = torch.empty((lib.b200lp_col_stats_parts(m), 2, c), dtype=torch.float32, device=x2d.device)
    with _timed("batchnorm", nbytes=4.0 * x2d.numel()):
        L.check(lib.b200lp_col_stats(L.ptr(x2d), L.ptr(part), m, c, L.stream_ptr()), "col_stats")
    return part


def bn_act(x, scale=None, shift=None, res=None, res_scale=None, res_shift=None, act=1, round_tf32=True, want_f32=True,
           want_split=False, want_mask=False):
    """act(x*scale+shift (+ res[*res_scale+res_shift])) over an (..., C) tensor; act 0 none / 1 relu / 2 relu6.
    Returns y (fp32), (y, y_split) or y_split with y_split (2, ...) bfloat16 (hi, lo) planes of the unrounded value;
    with want_mask additionally the uint8 (numel / 4,) activation bit mask that bn_bwd(mask_mode=4) reads (appended)."""
    lib = L.load()
    c = x.shape[-1]
    m = x.numel() // c
    y = torch.empty_like(x) if want_f32 else None
    ys = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device) if want_split else None
    mask = torch.empty((x.numel() // 4,), dtype=torch.uint8, device=x.device) if want_mask else None
    n_in = 1 + int(res is not None)
    with _timed("batchnorm", nbytes=4.0 * x.numel() * (n_in + int(want_f32) + int(want_split))):
        L.check(lib.b200lp_bn_act(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(res), L.ptr(res_scale), L.ptr(res_shift),
                                  L.ptr(y), L.ptr(ys, torch.bfloat16), m, c, int(act), int(round_tf32),
                                  L.ptr(mask, torch.uint8), L.stream_ptr()), "bn_act")
    out = (y, ys) if (want_f32 and want_split) else (y if want_f32 else ys)
    if want_mask:
        return (out + (mask,)) if isinstance(out, tuple) else (out, mask)
    return out


def bn_bwd(dy, x_raw, mean, rstd, gamma, scale=None, shift=None, mask_src=None, mask_mode=0, dgamma=None, dbeta=None,
           accumulate=False, batch_stats=True, round_tf32=False, want_dz=False):
    """BatchNorm(+activation) backward over (..., C) tensors; see include/b200lp.h.  Returns (dx, dgamma, dbeta, dz);
    dgamma / dbeta are the given buffers (added into when accumulate) or fresh tensors; dz only when want_dz."""
    lib = L.load()
    c = dy.shape[-1]
    m = dy.numel() // c
    ws = _ws(lib.b200lp_bn_bwd_workspace(m, c), dy.device)
    if dgamma is None:
        dgamma, dbeta, accumulate = (torch.empty(c, dtype=torch.float32, device=dy.device),
                                     torch.empty(c, dtype=torch.float32, device=dy.device), False)
    dx = torch.empty_like(dy)
    dz = torch.empty_like(dy) if want_dz else None
    n_streams = 3 + int(mask_mode == 1)
    with _timed("batchnorm", nbytes=4.0 * dy.numel() * (n_streams + int(want_dz))):   # ideal: dy, x read once, dx written
        L.check(lib.b200lp_bn_bwd(L.ptr(dy), L.ptr(mask_src, torch.uint8 if mask_mode == 4 else torch.float32), L.ptr(x_raw), L.ptr(mean), L.ptr(rstd), L.ptr(scale),
                                  L.ptr(shift), L.ptr(gamma), L.ptr(dgamma), L.ptr(dbeta), int(accumulate), L.ptr(dx),
                                  L.ptr(dz), L.ptr(ws), ws.numel() * 4, m, c, int(mask_mode), int(batch_stats),
                                  int(round_tf32), L.stream_ptr()), "bn_bwd")
    return dx, dgamma, dbeta, dz


def gconv3x3_fwd(x, w, in_scale=None, in_shift=None, stride=1, want_stats=False):
    """Grouped 3x3 conv (padding 1) on relu(x*in_scale+in_shift) (or x); x (N,H,W,C) raw, w (C, cpg, 3, 3)."""
    lib = L.load()
    n, h, wd, c = x.shape
    cpg = w.shape[1]
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device)
    part = None
    if want_stats:
        nparts = lib.b200lp_gconv3x3_parts(n, h, wd, c, cpg, stride)
        if nparts <= 0:
            raise L.B200lpError(f"gconv3x3: unsupported shape {tuple(x.shape)} cpg={cpg} stride={stride}")
        part = torch.empty((nparts, 2, c), dtype=torch.float32, device=x.device)
    with _timed("resnext_grouped", flops=2.0 * n * ho * wo * c * cpg * 9):
        L.check(lib.b200lp_gconv3x3_fwd(L.ptr(x), L.ptr(in_scale), L.ptr(in_shift), L.ptr(w), L.ptr(y), L.ptr(part), n, h,
                                        wd, c, cpg, stride, 0, L.stream_ptr()), "gconv3x3_fwd")
    return (y, part) if want_stats else y


def _pow2(v):
    return v >= 2 and (v & (v - 1)) == 0


def gconv_tensor_cores(n, h, w, c, cpg):
    """True when the grouped 3x3 gradients of this shape run on the tensor cores (block-diagonal 32-channel tiles):
    power-of-two planes, 32-aligned channels, groups inside one block.  B200LP_GCONV_CUDA_CORES=1 forces the FP32 kernels."""
    import os
    if os.environ.get("B200LP_GCONV_CUDA_CORES"):
        return False
    return _pow2(h) and _pow2(w) and c % 32 == 0 and cpg in (1, 2, 4, 8, 16, 32) and n * h * w >= 64 and \
        (h * w >= 64 or n % max(64 // (h * w), 1) == 0)


def zero_stuff2(x):
    """(N,H,W,C) -> (N,2H,2W,C) with x at the even positions, zeros elsewhere."""
    lib = L.load()
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=x.device)
    with _timed("resnext", nbytes=4.0 * x.numel() + 4.0 * out.numel()):
        L.check(lib.b200lp_zero_stuff2(L.ptr(x), L.ptr(out), n, h, w, c, L.stream_ptr()), "zero_stuff2")
    return out


def gconv3x3_wgrad_tc(x, dy, cpg, acc_into=None):
    """dw (C, cpg, 3, 3) (+)= grouped weight gradient of a stride-1 3x3 conv on the TF32 tensor cores; x, dy (N,H,W,C)."""
    lib = L.load()
    assert dy.shape == x.shape, (dy.shape, x.shape)
    n_alg = x.shape[0]
    x, dy = wgrad_batch_pad(x, dy)
    n, h, wd, c = x.shape
    nbytes = lib.b200lp_gconv3x3_wgrad_tc_workspace(n, h, wd, c)
    if nbytes <= 0:
        raise L.B200lpError(f"gconv3x3_wgrad_tc: unsupported shape {tuple(x.shape)}: {L.last_error()}")
    ws = _ws(nbytes, x.device)
    dw = acc_into if acc_into is not None else torch.empty((c, cpg, 3, 3), dtype=torch.float32, device=x.device)
    assert tuple(dw.shape) == (c, cpg, 3, 3) and dw.is_contiguous()
    a = L.WgradArgs()
    a.x = L.ptr(x); a.dy = L.ptr(dy); a.dw = L.ptr(dw); a.workspace = L.ptr(ws)
    a.workspace_bytes = ws.numel() * 4
    a.N, a.H, a.W, a.Cin, a.Cout = n, h, wd, c, c
    a.ksize = 3
    a.scale = 1.0
    a.grouped = cpg
    with _timed("resnext_grouped", flops=2.0 * n_alg * h * wd * c * cpg * 9):
        L.check(lib.b200lp_gconv3x3_wgrad_tc(byref(a), int(acc_into is not None), L.stream_ptr()), "gconv3x3_wgrad_tc")
    return dw


def gconv3x3_dgrad(dy, w, in_hw, stride=1, packed=None):
    """dy (N,Ho,Wo,C) -> dx (N,H,W,C) with (H, W) = in_hw.  `packed`: pack_gconv_weight(w, transpose=True) — then the
    gradient runs on the TF32 tensor cores (stride 2: on the zero-stuffed gradient `dy` must then already BE, i.e. the
    caller passes zero_stuff2(dy) and stride=1)."""
    lib = L.load()
    n, ho, wo, c = dy.shape
    h, wd = in_hw
    cpg = w.shape[1]
    if packed is not None:
        assert stride == 1 and (ho, wo) == (h, wd)
        return conv_fwd(dy, packed, 3, grouped=cpg)
    dx = torch.empty((n, h, wd, c), dtype=torch.float32, device=dy.device)
    with _timed("resnext_grouped", flops=2.0 * n * ho * wo * c * cpg * 9):
        L.check(lib.b200lp_gconv3x3_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), n, h, wd, c, cpg, stride, L.stream_ptr()),
                "gconv3x3_dgrad")
    return dx


def gconv3x3_wgrad(x, dy, cpg, in_scale=None, in_shift=None, stride=1, acc_into=None):
    """dw (C, cpg, 3, 3) = sum_pixels dy (x) relu(x*in_scale+in_shift); added into `acc_into` when given."""
    lib = L.load()
    n, h, wd, c = x.shape
    nbytes = lib.b200lp_gconv3x3_wgrad_workspace(n, h, wd, c, cpg, stride)
    if nbytes <= 0:
        raise L.B200lpError(f"gconv3x3_wgrad: unsupported shape {tuple(x.shape)} cpg={cpg} stride={stride}")
    ws = _ws(nbytes, x.device)
    dw = acc_into if acc_into is not None else torch.empty((c, cpg, 3, 3), dtype=torch.float32, device=x.device)
    assert tuple(dw.shape) == (c, cpg, 3, 3)
    with _timed("resnext_grouped", flops=2.0 * dy.numel() * cpg * 9):
        L.check(lib.b200lp_gconv3x3_wgrad(L.ptr(x), L.ptr(in_scale), L.ptr(in_shift), L.ptr(dy), L.ptr(dw),
                                          int(acc_into is not None), L.ptr(ws), ws.numel() * 4, n, h, wd, c, cpg, stride,
                                          L.stream_ptr()), "gconv3x3_wgrad")
    return dw


STEM_KP = 192      # 147 patch columns of the 7x7x3 stem, zero padded to a multiple of 64 (bf16x3 K rows)


def im2col7x7_s2(x_nchw, want_f32=True, want_split=True):
    """(N,3,H,W) -> patch matrix (N,Ho,Wo,STEM_KP): fp32 tf32-rounded and / or (2, ...) bf16 (hi, lo) planes."""
    lib = L.load()
    n, c, h, w = x_nchw.shape
    assert c == 3
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    col = torch.empty((n, ho, wo, STEM_KP), dtype=torch.float32, device=x_nchw.device) if want_f32 else None
    cols = torch.empty((2, n, ho, wo, STEM_KP), dtype=torch.bfloat16, device=x_nchw.device) if want_split else None
    with _timed("resnext", nbytes=4.0 * n * ho * wo * STEM_KP * (int(want_f32) + int(want_split))):
        L.check(lib.b200lp_im2col7x7_s2(L.ptr(x_nchw), L.ptr(col), L.ptr(cols, torch.bfloat16), n, h, w, STEM_KP,
                                        L.stream_ptr()), "im2col7x7_s2")
    if want_f32 and want_split:
        return col, cols
    return col if want_f32 else cols


def maxpool3x3s2_fwd(x, scale, shift, want_f32=True, want_split=False, want_idx=True, round_tf32=True):
    """maxpool3x3/s2/p1(relu(x*scale+shift)) -> (y | None, y_split | None, idx | None)."""
    lib = L.load()
    n, h, w, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, ho, wo, c), dtype=torch.float32, device=x.device) if want_f32 else None
    ys = torch.empty((2, n, ho, wo, c), dtype=torch.bfloat16, device=x.device) if want_split else None
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    with _timed("resnext", nbytes=4.0 * (x.numel() + n * ho * wo * c * (int(want_f32) + int(want_split)))):
        L.check(lib.b200lp_maxpool3x3s2_fwd(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(y), L.ptr(ys, torch.bfloat16),
                                            L.ptr(idx, torch.uint8), n, h, w, c, int(round_tf32), L.stream_ptr()),
                "maxpool3x3s2_fwd")
    return y, ys, idx


def maxpool3x3s2_bwd(dy, idx, in_hw):
    lib = L.load()
    n, ho, wo, c = dy.shape
    h, w = in_hw
    dx = torch.empty((n, h, w, c), dtype=torch.float32, device=dy.device)
    with _timed("resnext", nbytes=4.0 * (dx.numel() + dy.numel())):
        L.check(lib.b200lp_maxpool3x3s2_bwd(L.ptr(dy), L.ptr(idx, torch.uint8), L.ptr(dx), n, h, w, c, L.stream_ptr()),
                "maxpool3x3s2_bwd")
    return dx


def subsample2(x=None, x_split=None):
    """x (N,H,W,C) fp32 and / or x_split (2,N,H,W,C) bf16 -> the even pixels (N,H/2,W/2,C)."""
    lib = L.load()
    ref = x if x is not None else x_split[0]
    n, h, w, c = ref.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=ref.device) if x is not None else None
    ys = torch.empty((2, n, h // 2, w // 2, c), dtype=torch.bfloat16, device=ref.device) if x_split is not None else None
    with _timed("resnext", nbytes=2.0 * n * (h // 2) * (w // 2) * c * (4 * int(x is not None) + 4 * int(x_split is not None))):
        L.check(lib.b200lp_subsample2(L.ptr(x), L.ptr(x_split, torch.bfloat16), L.ptr(y), L.ptr(ys, torch.bfloat16), n,
                                      h // 2, w // 2, c, L.stream_ptr()), "subsample2")
    return y, ys


def scatter_add2(dsub, dx):
    """dx[:, ::2, ::2] += dsub (in place)."""
    lib = L.load()
    n, ho, wo, c = dsub.shape
    assert tuple(dx.shape) == (n, 2 * ho, 2 * wo, c)
    with _timed("resnext", nbytes=12.0 * dsub.numel()):
        L.check(lib.b200lp_scatter_add2(L.ptr(dsub), L.ptr(dx), n, ho, wo, c, L.stream_ptr()), "scatter_add2")
    return dx


def avgpool_fwd(x):
    lib = L.load()
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=torch.float32, device=x.device)
    with _timed("resnext", nbytes=4.0 * x.numel()):
        L.check(lib.b200lp_avgpool_fwd(L.ptr(x), L.ptr(y), n, h * w, c, L.stream_ptr()), "avgpool_fwd")
    return y


def avgpool_bwd(dy, hw):
    lib = L.load()
    n, c = dy.shape
    h, w = hw
    dx = torch.empty((n, h, w, c), dtype=torch.float32, device=dy.device)
    with _timed("resnext", nbytes=4.0 * dx.numel()):
        L.check(lib.b200lp_avgpool_bwd(L.ptr(dy), L.ptr(dx), n, h * w, c, L.stream_ptr()), "avgpool_bwd")
    return dx


def sgemm(a, b, trans_a=False, trans_b=False, acc_into=None, alpha=None, bias=None):
    """alpha * (a^T if trans_a else a) @ (b^T if trans_b else b) (+ bias per column) for small contiguous fp32 matrices
    (classifier / projector layers); alpha: 1-element device tensor; added into `acc_into` when given."""
    lib = L.load()
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    kb, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert k == kb, (a.shape, b.shape, trans_a, trans_b)
    sai, sak = (1, a.shape[1]) if trans_a else (a.shape[1], 1)
    sbk, sbj = (1, b.shape[1]) if trans_b else (b.shape[1], 1)
    out = acc_into if acc_into is not None else torch.empty((m, n), dtype=torch.float32, device=a.device)
    assert out.numel() == m * n
    need = lib.b200lp_sgemm_strided_workspace(m, n, k)
    ws = _ws(need, a.device) if need > 0 else None
    with _timed("dense_small", flops=2.0 * m * n * k):
        L.check(lib.b200lp_sgemm_strided(L.ptr(a), sai, sak, L.ptr(b), sbk, sbj, L.ptr(out), L.ptr(alpha), L.ptr(bias), m, n,
                                         k, int(acc_into is not None), L.ptr(ws), ws.numel() * 4 if ws is not None else 0,
                                         L.stream_ptr()), "sgemm_strided")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Scalar losses and the small dense pieces around them — csrc/losses.cu
# ---------------------------------------------------------------------------------------------------------------------
def dice_fwd(fake, real, weight):
    """fake (B, 1, H, W), real (B, CR, H, W) -> (loss (1,), sums (3,)) of criterions/dice.py:30-34."""
    lib = L.load()
    b, cr = real.shape[0], real.shape[1]
    hw = real.shape[2] * real.shape[3]
    assert fake.numel() == b * hw
    ws = _ws(lib.b200lp_dice_workspace(b, hw), fake.device)
    sums = torch.empty(3, dtype=torch.float32, device=fake.device)
    loss = torch.empty(1, dtype=torch.float32, device=fake.device)
    with _timed("losses", nbytes=4.0 * (fake.numel() + real.numel())):
        L.check(lib.b200lp_dice_fwd(L.ptr(fake), L.ptr(real), c_float(weight), L.ptr(sums), L.ptr(loss), L.ptr(ws),
                                    ws.numel() * 4, b, cr, hw, L.stream_ptr()), "dice_fwd")
    return loss, sums


def dice_bwd(fake, real, sums, grad, weight):
    lib = L.load()
    b, cr = real.shape[0], real.shape[1]
    hw = real.shape[2] * real.shape[3]
    d_fake = torch.empty_like(fake)
    with _timed("losses", nbytes=4.0 * (2 * fake.numel() + real.numel())):
        L.check(lib.b200lp_dice_bwd(L.ptr(fake), L.ptr(real), L.ptr(sums), L.ptr(grad), c_float(weight), L.ptr(d_fake), b, cr,
                                    hw, L.stream_ptr()), "dice_bwd")
    return d_fake


def adversarial_fwd(fake_g, fake_d, real, relativistic=0):
    """(B,) score vectors -> out (2,) = (loss_G, loss_D)."""
    lib = L.load()
    out = torch.empty(2, dtype=torch.float32, device=real.device)
    L.check(lib.b200lp_adversarial_fwd(L.ptr(fake_g), L.ptr(fake_d), L.ptr(real), L.ptr(out), real.numel(), relativistic,
                                       L.stream_ptr()), "adversarial_fwd")
    return out


def adversarial_bwd(fake_d, real, grad_g, grad_d, need_g=True, need_d=True):
    """-> (d_fake_g, d_fake_d, d_real) for the `gan` type; grad_g / grad_d: 1-element device tensors or None."""
    lib = L.load()
    b = real.numel()
    dg = torch.empty_like(real) if need_g and grad_g is not None else None
    dd = torch.empty_like(real) if need_d and grad_d is not None else None
    dr = torch.empty_like(real) if need_d and grad_d is not None else None
    L.check(lib.b200lp_adversarial_bwd(L.ptr(fake_d), L.ptr(real), L.ptr(grad_g), L.ptr(grad_d), L.ptr(dg), L.ptr(dd),
                                       L.ptr(dr), b, L.stream_ptr()), "adversarial_bwd")
    return dg, dd, dr


def crop_bilinear_fwd(x, boxes, out_hw=None):
    """x (B, C, H, W), boxes (B, 4) [t, b, l, r] pixels (device) -> (B, C, OH, OW): affine_grid + grid_sample(bilinear,
    reflection, align_corners=False)."""
    lib = L.load()
    b, c, h, w = x.shape
    oh, ow = out_hw or (h, w)
    y = torch.empty((b, c, oh, ow), dtype=torch.float32, device=x.device)
    with _timed("losses", nbytes=4.0 * (x.numel() + y.numel())):
        L.check(lib.b200lp_crop_bilinear_fwd(L.ptr(x), L.ptr(boxes), L.ptr(y), b, c, h, w, oh, ow, L.stream_ptr()),
                "crop_bilinear_fwd")
    return y


def crop_bilinear_bwd(dy, boxes, in_hw):
    lib = L.load()
    b, c, oh, ow = dy.shape
    h, w = in_hw
    dx = torch.empty((b, c, h, w), dtype=torch.float32, device=dy.device)
    with _timed("losses", nbytes=4.0 * (dx.numel() + dy.numel())):
        L.check(lib.b200lp_crop_bilinear_bwd(L.ptr(dy), L.ptr(boxes), L.ptr(dx), b, c, h, w, oh, ow, L.stream_ptr()),
                "crop_bilinear_bwd")
    return dx


def disc_head_fwd(feat, embed, w, inv_sigma, bias):
    """feat (B, H, W, C) NHWC raw -> (score (B,), o (B, C))."""
    lib = L.load()
    b, h, wd, c = feat.shape
    o = torch.empty((b, c), dtype=torch.float32, device=feat.device)
    score = torch.empty(b, dtype=torch.float32, device=feat.device)
    L.check(lib.b200lp_disc_head_fwd(L.ptr(feat), L.ptr(embed), L.ptr(w), L.ptr(inv_sigma), L.ptr(bias), L.ptr(o),
                                     L.ptr(score), b, h * wd, c, L.stream_ptr()), "disc_head_fwd")
    return score, o


def disc_head_bwd(feat, embed, w, inv_sigma, o, grad, need_feat=True, need_embed=True, need_params=True, dw_acc=None,
                  db_acc=None):
    """-> (d_feat, d_embed, dw, ds, dbias); dw / dbias are added into dw_acc / db_acc when given."""
    lib = L.load()
    b, h, wd, c = feat.shape
    d_feat = torch.empty_like(feat) if need_feat else None
    d_embed = torch.empty((b, c), dtype=torch.float32, device=feat.device) if (need_embed and embed is not None and need_feat) else None
    dw = ds = db = None
    acc = 0
    if need_params:
        acc = int(dw_acc is not None)
        dw = dw_acc if dw_acc is not None else torch.empty(c, dtype=torch.float32, device=feat.device)
        db = db_acc if db_acc is not None else torch.empty(1, dtype=torch.float32, device=feat.device)
        ds = torch.empty(1, dtype=torch.float32, device=feat.device)
    L.check(lib.b200lp_disc_head_bwd(L.ptr(feat), L.ptr(embed), L.ptr(w), L.ptr(inv_sigma), L.ptr(o), L.ptr(grad),
                                     L.ptr(d_feat), L.ptr(d_embed), L.ptr(dw), L.ptr(ds), L.ptr(db), acc, b, h * wd, c,
                                     L.stream_ptr()), "disc_head_bwd")
    return d_feat, d_embed, dw, ds, db


# ---------------------------------------------------------------------------------------------------------------------
# Pose encoder (MobileNetV2) backward — csrc/mobilenet_bwd.cu
# ---------------------------------------------------------------------------------------------------------------------
def transpose2d(src):
    lib = L.load()
    r, c = src.shape
    dst = torch.empty((c, r), dtype=torch.float32, device=src.device)
    L.check(lib.b200lp_transpose2d(L.ptr(src), L.ptr(dst), r, c, L.stream_ptr()), "transpose2d")
    return dst


def pw_wgrad(dy2d, x2d, in_scale=None, in_shift=None, in_relu6=False, acc_into=None):
    """dw (Cout, Cin) = dy^T f(x), f = producer BatchNorm (+ReLU6) on load; added into `acc_into` when given."""
    lib = L.load()
    m, cout = dy2d.shape
    cin = x2d.shape[1]
    ws = _ws(lib.b200lp_pw_wgrad_workspace(m, cin, cout), dy2d.device)
    dw = acc_into if acc_into is not None else torch.empty((cout, cin), dtype=torch.float32, device=dy2d.device)
    assert dw.numel() == cout * cin
    with _timed("pose_encoder", flops=2.0 * m * cin * cout):
        L.check(lib.b200lp_pw_wgrad(L.ptr(dy2d), L.ptr(x2d), L.ptr(in_scale), L.ptr(in_shift), int(in_relu6), L.ptr(dw),
                                    int(acc_into is not None), L.ptr(ws), ws.numel() * 4, m, cin, cout, L.stream_ptr()),
                "pw_wgrad")
    return dw


def dw_dgrad(dy, w, in_hw, stride):
    lib = L.load()
    n, ho, wo, c = dy.shape
    h, wd = in_hw
    dx = torch.empty((n, h, wd, c), dtype=torch.float32, device=dy.device)
    with _timed("pose_encoder", nbytes=4.0 * (dy.numel() + dx.numel())):
        L.check(lib.b200lp_dw_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), n, h, wd, c, stride, L.stream_ptr()), "dw_dgrad")
    return dx


def dw_wgrad(x, dy, in_scale, in_shift, stride, acc_into=None):
    """dw (C, 1, 3, 3) = sum_p dy * relu6(x*scale+shift) at the 9 taps; added into `acc_into` when given."""
    lib = L.load()
    n, h, wd, c = x.shape
    ws = _ws(lib.b200lp_dw_wgrad_workspace(n, h, wd, c, stride), x.device)
    dw = acc_into if acc_into is not None else torch.empty((c, 1, 3, 3), dtype=torch.float32, device=x.device)
    assert dw.numel() == c * 9
    with _timed("pose_encoder", nbytes=4.0 * (x.numel() + dy.numel())):
        L.check(lib.b200lp_dw_wgrad(L.ptr(x), L.ptr(in_scale), L.ptr(in_shift), L.ptr(dy), L.ptr(dw),
                                    int(acc_into is not None), L.ptr(ws), ws.numel() * 4, n, h, wd, c, stride,
                                    L.stream_ptr()), "dw_wgrad")
    return dw


def mbv2_stem_wgrad(x_nchw, dy, acc_into=None):
    lib = L.load()
    n, _, h, w = x_nchw.shape
    ws = _ws(lib.b200lp_mbv2_stem_wgrad_workspace(n, h, w), dy.device)
    dw = acc_into if acc_into is not None else torch.empty((32, 3, 3, 3), dtype=torch.float32, device=dy.device)
    with _timed("pose_encoder", nbytes=4.0 * (x_nchw.numel() + dy.numel())):
        L.check(lib.b200lp_mbv2_stem_wgrad(L.ptr(x_nchw), L.ptr(dy), L.ptr(dw), int(acc_into is not None), L.ptr(ws),
                                           ws.numel() * 4, n, h, w, L.stream_ptr()), "mbv2_stem_wgrad")
    return dw
