"""torch.autograd.Function wrappers around the C-ABI kernels (b200lp.kernels).

torch.autograd is used as the graph/plumbing layer only (the reference's runner calls `loss.backward()` twice with
`retain_graph`, runners/holycow.py:239-252); every forward and backward computation below is a libb200lp kernel.
Activations are NHWC (N, H, W, C) float32 CUDA tensors.
"""
import weakref

import torch

from . import kernels as K


def _dot(a, b):
    return (a * b).sum()


_GENERATION = [0]
_PTR_GENERATION = {}


def bump_generation(params=None):
    """Call after weights were modified through raw pointers (fused optimizer / EMA kernels, CUDA-graph replays):
    those writes do not touch torch's per-tensor version counters, so the cached packed copies are invalidated here —
    those of `params` (keyed by storage pointer) or, without arguments, all of them."""
    if params is None:
        _GENERATION[0] += 1
        return
    for p in params:
        ptr = p.data_ptr()
        _PTR_GENERATION[ptr] = _PTR_GENERATION.get(ptr, 0) + 1


def _stamp(weight):
    ptr = weight.data_ptr()
    return (_GENERATION[0], _PTR_GENERATION.get(ptr, 0), weight._version, ptr)


class PackCache:
    """Packed (tensor-core layout) copies of one weight, keyed by (transpose, precision) and validated by the weight
    tensor's autograd version counter plus a global generation (see bump_generation): a weight is re-packed once per
    update instead of once per forward call (the three discriminator passes of a step and the backward passes share
    the copies).  Lives on the owning module; deep copies start empty.  Live caches are tracked so that
    `repack_weights` can refresh every copy of a network with one launch after its optimizer step."""

    _live = weakref.WeakSet()

    def __init__(self):
        self.entries = {}
        PackCache._live.add(self)

    def __deepcopy__(self, memo):
        return PackCache()

    def get(self, weight, transpose, precision):
        key = (transpose, precision)
        hit = self.entries.get(key)
        stamp = _stamp(weight)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        # a stale copy is refreshed IN PLACE: its address is baked into captured CUDA graphs and repack tables
        reuse = hit[1] if hit is not None and hit[2] == tuple(weight.shape) and hit[1].device == weight.device else None
        packed = K.pack_conv_weight(weight.detach(), None, transpose=transpose, precision=precision, out=reuse)
        self.entries[key] = (stamp, packed, tuple(weight.shape))
        return packed


_REPACK_PLANS = {}


def repack_weights(params, owner=None):
    """Refresh, with ONE multi-tensor launch, every packed copy that exists of the weights in `params` (call right after
    the optimizer kernel that rewrote them, after bump_generation).  Copies are created lazily by PackCache.get on first
    use; from the second step on all of them are known and this replaces ~50 single-tensor pack launches per network.
    `owner`: key under which the device-side table is cached (the optimizer)."""
    by_ptr = {p.data_ptr(): p for p in params if p.is_cuda}
    rows, hits = [], []
    for cache in list(PackCache._live):
        for key, (stamp, packed, shape) in cache.entries.items():
            w = by_ptr.get(stamp[3])
            if w is None or tuple(w.shape) != shape:
                continue
            co, ci, kh, kw = shape
            rows.append((stamp[3], packed.data_ptr(), co, ci, kh * kw, int(key[0]), int(key[1]), w.numel()))
            hits.append((cache, key, packed, shape, w))
    if not rows:
        return 0
    sig = tuple(rows)
    plan = _REPACK_PLANS.get(owner)
    if plan is None or plan["sig"] != sig:
        if torch.cuda.is_current_stream_capturing():
            # the table upload is not capturable: refresh copy by copy (same kernels as the lazy path)
            for cache, key, packed, shape, w in hits:
                K.pack_conv_weight(w.detach(), None, transpose=key[0], precision=key[1], out=packed)
                cache.entries[key] = (_stamp(w), packed, shape)
            return len(rows)
        plan = K.pack_plan(rows, hits[0][4].device)
        plan["sig"] = sig
        _REPACK_PLANS[owner] = plan
    K.pack_conv_weight_multi(plan)
    for cache, key, packed, shape, w in hits:
        cache.entries[key] = (_stamp(w), packed, shape)
    return len(rows)


# ----------------------------------------------------------------------------------------------------------------
# Gradient sinks: while `direct_grads(...)` is active, the backward of the tensor-core convolutions adds a parameter's
# gradient straight into its `.grad` buffer (a view into the runner's flat gradient bucket) from inside the kernel
# that produces it, and hands autograd `None` for that input — instead of: temporary dW tensor -> spectral-norm fix
# kernel pair -> AccumulateGrad's `add_`.  Keyed by the parameter's storage pointer (saved tensors / detached aliases
# of a parameter share it).
# ----------------------------------------------------------------------------------------------------------------
_SINKS = {}


class direct_grads:
    """Context manager: `sinks` = {param.data_ptr(): contiguous fp32 gradient buffer shaped like the parameter}, already
    zeroed (or holding the gradient accumulated so far); autograd hooks of those parameters do not fire."""

    def __init__(self, sinks):
        self.sinks = sinks or {}

    def __enter__(self):
        self.old = dict(_SINKS)
        _SINKS.update(self.sinks)
        return self

    def __exit__(self, *exc):
        _SINKS.clear()
        _SINKS.update(self.old)
        return False


def _sink(t):
    if t is None or not _SINKS:
        return None
    g = _SINKS.get(t.data_ptr())
    if g is None or g.shape != t.shape or g.device != t.device:
        return None
    return g


def _round_gradients():
    """Gradients that only feed tf32 MMAs are stored rounded to nearest (the MMA would truncate: a 2^-12 relative bias per
    layer that compounds along the chain).  B200LP_DISC_NODE_TRUNCATE=1 (A/B tests) keeps the unrounded values."""
    import os
    return not os.environ.get("B200LP_DISC_NODE_TRUNCATE")


def _packed(weight, cache, transpose, precision=K.TF32):
    if cache is not None:
        return cache.get(weight, transpose, precision)
    return K.pack_conv_weight(weight.detach(), None, transpose=transpose, precision=precision)


def _conv_backward(ctx_ksize, x_f32, weight_orig, inv_sigma, dy, need_x, need_w, need_s, cache=None, sn=None):
    """Shared backward of the tensor-core convolutions (TF32 data / weight gradient kernels).
    y = s * conv(x, W):  dx = s * conv_T(dy, W);  dW = s * wgrad(x, dy);  ds = <wgrad(x, dy), W>.
    `sn` = (u, v) snapshots of the spectral-norm vectors behind s = 1/sigma: then s carries no autograd edge and the
    full gradient  dW = s*G - s^2 <G, W> u v^T   is produced here by one fused kernel pair."""
    dx = dw = ds = None
    if need_x:
        wpt = _packed(weight_orig, cache, True)
        dx = K.conv_fwd(dy, wpt, ctx_ksize, scale=inv_sigma)
    sink = _sink(weight_orig) if need_w and (sn is not None or inv_sigma is None) else None
    if sink is not None:
        # gradient goes straight into the parameter's .grad buffer: wgrad -> tiled reduce (+= s*G, <G,W> partials)
        # -> rank-1 spectral-norm term; autograd gets None for this input
        if sn is not None:
            K.conv_wgrad_sn_acc(x_f32, dy, ctx_ksize, sink, weight_orig, inv_sigma, sn[0], sn[1])
        else:
            K.conv_wgrad_sn_acc(x_f32, dy, ctx_ksize, sink)
    elif need_w or (need_s and inv_sigma is not None and sn is None):
        g = K.conv_wgrad(x_f32, dy, ctx_ksize)
        if sn is not None:
            dw = K.sn_wgrad_fix(g, weight_orig, inv_sigma, sn[0], sn[1])
        elif inv_sigma is not None:
            if need_s:
                ds = _dot(g, weight_orig).reshape(inv_sigma.shape)
            if need_w:
                dw = g * inv_sigma
        else:
            dw = g
    return dx, dw, ds


class InvSigmaFn(torch.autograd.Function):
    """Gives a batched-kernel 1/sigma (b200lp_sn_sigma_multi: no autograd history) its edge to `weight_orig`:
    sigma = u^T W v with u, v constants (torch's SpectralNorm.compute_weight)  =>  d(1/sigma)/dW = -(1/sigma)^2 u v^T.
    For the spectral-normalised layers whose consumers differentiate through plain autograd (linear / embedding layers,
    the Cin=3 stems, the generator tail); the tensor-core convs fuse this term into their weight-gradient kernels."""

    @staticmethod
    def forward(ctx, weight_orig, inv_sigma, u, v):
        ctx.save_for_backward(inv_sigma, u, v)
        ctx.shape = weight_orig.shape
        return inv_sigma.clone()

    @staticmethod
    def backward(ctx, ds):
        inv_sigma, u, v = ctx.saved_tensors
        coef = -(ds.reshape(1) * inv_sigma * inv_sigma)
        return ((u * coef).unsqueeze(1) * v.unsqueeze(0)).reshape(ctx.shape), None, None, None


def inv_sigma_edge(weight_orig, inv_sigma, u, v):
    return InvSigmaFn.apply(weight_orig, inv_sigma, u, v)


class Conv2dFn(torch.autograd.Function):
    """y = epilogue(conv_k(x, weight_orig * inv_sigma)).  Replaces nn.Conv2d under spectral_norm
    (generators/common/blocks.py:78-100, discriminators/no_landmarks.py:54-66) and its autograd backward.

    Gradients: dx (same tensor-core kernel, transposed packing), d(weight_orig) = G*inv_sigma and
    d(inv_sigma) = <G, weight_orig> with G the tensor-core weight gradient; sigma's own dependence on weight_orig
    (u^T W v) is left to autograd through `inv_sigma`.
    """

    @staticmethod
    def forward(ctx, x, weight_orig, inv_sigma, bias, residual, ksize, residual_mode, relu, round_out, x_split,
                emit_split, cache, sn):
        """`x_split` (optional, non-differentiable): the (hi, lo) bf16 planes of x — when given, the forward runs in
        bf16x3 precision on them; gradients still flow to `x`.  `emit_split`: also return the (hi, lo) planes of y."""
        if x_split is not None:
            wp = _packed(weight_orig, cache, False, K.BF16X3)
            src = x_split
        else:
            wp = _packed(weight_orig, cache, False)
            src = x
        out = K.conv_fwd(src, wp, ksize, bias=bias, residual=residual, residual_mode=residual_mode, relu=relu,
                         round_tf32=round_out, emit_split=emit_split, scale=inv_sigma)
        y, y_split = out if emit_split else (out, None)
        ctx.ksize, ctx.residual_mode, ctx.relu = ksize, residual_mode, relu
        ctx.cache, ctx.sn = cache, sn
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias.detach() if bias is not None else None      # alias only: identifies the gradient sink
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, weight_orig, inv_sigma, y if relu else None)
        if emit_split:
            ctx.mark_non_differentiable(y_split)
            return y, y_split
        return y

    @staticmethod
    def backward(ctx, dy, *_unused):
        x, weight_orig, inv_sigma, y = ctx.saved_tensors
        dy = dy.contiguous()
        need_x, need_w, need_s, need_b, need_r = ctx.needs_input_grad[:5]
        bias_done = False
        if ctx.relu:
            bsink = _sink(ctx.bias_ref) if (ctx.has_bias and need_b) else None
            if bsink is not None and y.shape[-1] % 4 == 0 and y.shape[-1] <= 1024:
                # ReLU mask, bias column sums and tf32 rounding of the gradient operand in one pass
                dy = K.relu_bwd_fused(y, dy, bias_a=bsink, round_tf32=_round_gradients())
                bias_done = True
            else:
                dy = K.relu_bwd(y, dy)
        dx, dw, ds = _conv_backward(ctx.ksize, x, weight_orig, inv_sigma, dy, need_x, need_w, need_s, ctx.cache, ctx.sn)
        db = dr = None
        if ctx.has_bias and need_b and not bias_done:
            bsink = _sink(ctx.bias_ref)
            if bsink is not None:
                K.bias_grad(dy, acc_into=bsink)
            else:
                db = K.bias_grad(dy)
        if ctx.has_res and need_r:
            dr = dy if ctx.residual_mode == 1 else K.upsample2_bwd(dy)
        return dx, dw, ds, db, dr, None, None, None, None, None, None, None, None


def conv2d(x, weight_orig, inv_sigma=None, bias=None, residual=None, ksize=3, residual_mode=0, relu=False,
           round_out=False, x_split=None, emit_split=False, cache=None, sn=None):
    if residual is None:
        residual_mode = 0
    return Conv2dFn.apply(x, weight_orig, inv_sigma, bias, residual, ksize, residual_mode, relu, round_out, x_split,
                          emit_split, cache, sn)


class AdaINConvFn(torch.autograd.Function):
    """One generator half-block as ONE autograd node, bf16x3 precision:
        y = conv3x3( relu( instance_norm(x)*gamma + beta ) [nearest 2x] ; W/sigma ) (+ residual)
    (generators/common/blocks.py:70-88 + AdaptiveNorm2d :18-26).  Kernels: in_stats -> adain_relu (writes the
    (hi, lo) bf16 operand planes, plus a tf32 fp32 copy when a backward pass will need it for the weight gradient)
    -> pack (hi, lo) -> tcgen05 bf16x3 implicit GEMM with the residual / skip-upsample epilogue.
    Backward: TF32 data- and weight-gradient kernels, then the AdaIN backward kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, weight_orig, inv_sigma, residual, eps, upsample2, residual_mode, emit_split,
                cache, sn):
        need_bwd = any(ctx.needs_input_grad)      # False under torch.no_grad() (drive.py, EMA forward)
        # statistics, then AdaIN + ReLU (+2x) writing the operand planes (K.adain_stats_apply)
        if need_bwd:
            mean, rstd, (a_f32, a_split) = K.adain_stats_apply(x, gamma, beta, eps, upsample2=upsample2, round_tf32=True,
                                                               want_f32=True, want_split=True)
        else:
            mean, rstd, a_split = K.adain_stats_apply(x, gamma, beta, eps, upsample2=upsample2, want_f32=False,
                                                      want_split=True)
            a_f32 = None
        wp = _packed(weight_orig, cache, False, K.BF16X3)
        out = K.conv_fwd(a_split, wp, 3, residual=residual, residual_mode=residual_mode, emit_split=emit_split,
                         scale=inv_sigma)
        y, y_split = out if emit_split else (out, None)
        ctx.upsample2, ctx.residual_mode = upsample2, residual_mode
        ctx.cache, ctx.sn = cache, sn
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, mean, rstd, gamma, beta, weight_orig, inv_sigma, a_f32)
        if emit_split:
            ctx.mark_non_differentiable(y_split)
            return y, y_split
        return y

    @staticmethod
    def backward(ctx, dy, *_unused):
        x, mean, rstd, gamma, beta, weight_orig, inv_sigma, a_f32 = ctx.saved_tensors
        dy = dy.contiguous()
        need_x, need_g, need_b, need_w, need_s, need_r = ctx.needs_input_grad[:6]
        da, dw, ds = _conv_backward(3, a_f32, weight_orig, inv_sigma, dy, need_x or need_g or need_b, need_w, need_s,
                                    ctx.cache, ctx.sn)
        dx = dgm = dbt = None
        if da is not None:
            dx, dgm, dbt = K.adain_relu_bwd(x, mean, rstd, gamma, beta, da, upsample2=ctx.upsample2)
        dr = None
        if ctx.has_res and need_r:
            dr = dy if ctx.residual_mode == 1 else K.upsample2_bwd(dy)
        return dx, dgm, dbt, dw, ds, dr, None, None, None, None, None, None


class AdaResBlockFn(torch.autograd.Function):
    """A whole generator block as ONE autograd node, bf16x3 precision (generators/common/blocks.py:47-111 with
    norm_layer='adain'; reference Sequential AdaIN -> ReLU -> [2x] -> conv -> AdaIN -> ReLU -> conv, + skip):
        y1  = conv3x3( relu(adain(x; g0, b0)) [2x] ; W0/s0 )
        s   = conv1x1(x; Ws/ss) + bs at the LOW resolution, or x
        out = conv3x3( relu(adain(y1; g1, b1)) ; W1/s1 ) + s [nearest 2x in the epilogue]
    Forward = the kernels of two AdaINConvFn + Conv2dFn.  The hand-scheduled backward merges the block input's two
    gradients (main branch through the first AdaIN, skip branch) inside the AdaIN-backward apply pass instead of an
    `at::add`, and stores every gradient that only feeds tf32 MMAs rounded to tf32 instead of letting the MMA truncate
    (the generator's first blocks sit behind 16 gradient convs: truncation alone shrank their gradient norms by 0.5 %).
    convs: c0 = (w, s, cache, sn), c1 likewise, sk = (w, s, b, cache, sn) or None."""

    @staticmethod
    def forward(ctx, x, g0, b0, g1, b1, x_split, convs, upsample, emit_split, eps, *params):
        ctx.set_materialize_grads(False)
        w0, s0, c0, _ = convs["c0"]
        w1, s1, c1, _ = convs["c1"]
        need_bwd = any(ctx.needs_input_grad)
        # each half-block: statistics (2 launches), AdaIN + ReLU [+2x] -> operand planes, the tensor-core conv
        if need_bwd:
            mean0, rstd0, (a0_f32, a0_split) = K.adain_stats_apply(x, g0, b0, eps, upsample2=upsample, round_tf32=True,
                                                                   want_f32=True, want_split=True)
        else:
            mean0, rstd0, a0_split = K.adain_stats_apply(x, g0, b0, eps, upsample2=upsample, want_f32=False, want_split=True)
            a0_f32 = None
        y1 = K.conv_fwd(a0_split, _packed(w0, c0, False, K.BF16X3), 3, scale=s0)
        del a0_split
        if convs["sk"] is not None:
            ws, ss, bs, cs, _ = convs["sk"]
            if x_split is not None:
                s = K.conv_fwd(x_split, _packed(ws, cs, False, K.BF16X3), 1, bias=bs, scale=ss)
            else:
                s = K.conv_fwd(x, _packed(ws, cs, False), 1, bias=bs, scale=ss)
            mode = 2 if upsample else 1
        else:
            s, mode = x, 1
        if need_bwd:
            mean1, rstd1, (a1_f32, a1_split) = K.adain_stats_apply(y1, g1, b1, eps, round_tf32=True, want_f32=True,
                                                                   want_split=True)
        else:
            mean1, rstd1, a1_split = K.adain_stats_apply(y1, g1, b1, eps, want_f32=False, want_split=True)
            a1_f32 = None
        out = K.conv_fwd(a1_split, _packed(w1, c1, False, K.BF16X3), 3, residual=s, residual_mode=mode,
                         emit_split=emit_split, scale=s1)
        y, y_split = out if emit_split else (out, None)
        ctx.convs, ctx.upsample, ctx.mode = convs, upsample, mode
        ctx.save_for_backward(x, mean0, rstd0, g0, b0, a0_f32, y1, mean1, rstd1, g1, b1, a1_f32)
        if emit_split:
            ctx.mark_non_differentiable(y_split)
            return y, y_split
        return y

    @staticmethod
    def backward(ctx, dy, *_unused):
        x, mean0, rstd0, g0, b0, a0_f32, y1, mean1, rstd1, g1, b1, a1_f32 = ctx.saved_tensors
        convs = ctx.convs
        w0, s0, c0, sn0 = convs["c0"]
        w1, s1, c1, sn1 = convs["c1"]
        wants = ctx.needs_input_grad[10:]        # params: w0, w1[, ws, bs]
        pgrads = [None] * len(wants)
        dy = dy.contiguous()

        def wgrad(slot, ksize, xin, w, s, cache, sn, d):
            if wants[slot]:
                _, dw, _ = _conv_backward(ksize, xin, w, s, d, False, True, False, cache, sn)
                pgrads[slot] = dw

        # second half-block
        da1 = K.conv_fwd(dy, _packed(w1, c1, True), 3, scale=s1)
        wgrad(1, 3, a1_f32, w1, s1, c1, sn1, dy)
        dy1, dg1, db1 = K.adain_relu_bwd(y1, mean1, rstd1, g1, b1, da1, round_tf32=True)
        del da1
        # skip branch: gradient w.r.t. the block input
        dr = dy if ctx.mode == 1 else K.upsample2_bwd(dy)
        if convs["sk"] is not None:
            ws, ss, bs, cs, sns = convs["sk"]
            d_skip = K.conv_fwd(dr, _packed(ws, cs, True), 1, scale=ss)
            wgrad(2, 1, x, ws, ss, cs, sns, dr)
            if bs is not None and wants[3]:
                bsink = _sink(bs)
                if bsink is not None:
                    K.bias_grad(dr, acc_into=bsink)
                else:
                    pgrads[3] = K.bias_grad(dr)
        else:
            d_skip = dr
        # first half-block; the skip gradient joins in the AdaIN-backward apply pass
        da0 = K.conv_fwd(dy1, _packed(w0, c0, True), 3, scale=s0)
        wgrad(0, 3, a0_f32, w0, s0, c0, sn0, dy1)
        dx, dg0, db0 = K.adain_relu_bwd(x, mean0, rstd0, g0, b0, da0, upsample2=ctx.upsample, add=d_skip, round_tf32=True)
        return (dx, dg0, db0, dg1, db1, None, None, None, None, None, *pgrads)


def ada_res_block(x, g0, b0, g1, b1, x_split, convs, upsample, emit_split, eps=1e-4):
    params = [convs["c0"][0], convs["c1"][0]]
    if convs["sk"] is not None:
        params += [convs["sk"][0], convs["sk"][2]]
    return AdaResBlockFn.apply(x, g0, b0, g1, b1, x_split, convs, upsample, emit_split, eps, *params)


def adain_conv(x, gamma, beta, weight_orig, inv_sigma, residual=None, residual_mode=0, eps=1e-4, upsample2=False,
               emit_split=False, cache=None, sn=None):
    if residual is None:
        residual_mode = 0
    return AdaINConvFn.apply(x, gamma, beta, weight_orig, inv_sigma, residual, eps, upsample2, residual_mode,
                             emit_split, cache, sn)


class SplitAffineFn(torch.autograd.Function):
    """The projector output (B, sum 2C) cut into the per-AdaIN (beta, gamma) column blocks as ONE autograd node
    (generators/vector_pose_unsupervised_segmentation_noBottleneck.py:108-125, `assign_affine_params`).  With plain
    slicing every one of the 34 slices gets its own backward: a zero-filled (B, 13056) tensor, a slice copy and an
    accumulation — ~100 tiny launches per generator backward; here the 34 incoming gradients are concatenated once.
    Outputs are column views (row stride = sum 2C) of one private copy: the kernels read them through
    (pointer, affine_stride)."""

    @staticmethod
    def forward(ctx, affine, sizes):
        base = affine.detach().clone(memory_format=torch.contiguous_format)
        ctx.sizes = tuple(sizes)
        outs, off = [], 0
        for c in ctx.sizes:
            outs.append(base[:, off:off + c])              # beta  (bias of the AdaIN)
            outs.append(base[:, off + c:off + 2 * c])      # gamma (weight)
            off += 2 * c
        assert off == affine.shape[1], (off, affine.shape)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        parts = []
        for i, g in enumerate(grads):
            c = ctx.sizes[i // 2]
            parts.append(g if g is not None else ref.new_zeros((ref.shape[0], c)))
        return torch.cat(parts, dim=1), None


def split_affine(affine, sizes):
    """-> [(gamma_i, beta_i)] for the AdaIN layers in `sizes` order."""
    outs = SplitAffineFn.apply(affine, tuple(sizes))
    return [(outs[2 * i + 1], outs[2 * i]) for i in range(len(sizes))]


class AdaINReLUFn(torch.autograd.Function):
    """relu(instance_norm(x) * gamma + beta) [nearest 2x] [tf32].  Replaces AdaptiveNorm2d.forward + ReLU + Upsample
    (generators/common/blocks.py:18-26,73,75) and their backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, upsample2, round_out):
        mean, rstd, y = K.adain_stats_apply(x, gamma, beta, eps, upsample2=upsample2, round_tf32=round_out)
        ctx.upsample2 = upsample2
        ctx.save_for_backward(x, mean, rstd, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma, beta = ctx.saved_tensors
        dx, dg, db = K.adain_relu_bwd(x, mean, rstd, gamma, beta, dy.contiguous(), upsample2=ctx.upsample2)
        return dx, dg, db, None, None, None


def adain_relu(x, gamma, beta, eps=1e-4, upsample2=False, round_out=True):
    return AdaINReLUFn.apply(x, gamma, beta, eps, upsample2, round_out)


class GenTailFn(torch.autograd.Function):
    """conv3x3(Cin->4)+bias -> tanh -> rgb*segm composition (generator :84-88,165-181), NHWC in, NCHW images out."""

    @staticmethod
    def forward(ctx, x, weight_orig, inv_sigma, bias):
        rgbs, segm, t = K.gen_tail_fwd(x, weight_orig, inv_sigma, bias)
        ctx.save_for_backward(x, weight_orig, inv_sigma, t)
        return rgbs, segm

    @staticmethod
    def backward(ctx, d_rgbs, d_segm):
        x, weight_orig, inv_sigma, t = ctx.saved_tensors
        d_rgbs = d_rgbs.contiguous() if d_rgbs is not None else None
        d_segm = d_segm.contiguous() if d_segm is not None else None
        need_x, need_w, need_s, need_b = ctx.needs_input_grad
        dx, g, db = K.gen_tail_bwd(x, t, weight_orig, inv_sigma, d_rgbs, d_segm, need_dx=need_x,
                                   need_dw=(need_w or need_s or need_b))
        dw = ds = None
        if g is not None:
            if need_s:
                ds = _dot(g, weight_orig).reshape(inv_sigma.shape)
            if need_w:
                dw = g * inv_sigma
        return dx, dw, ds, (db if need_b else None)


def gen_tail(x, weight_orig, inv_sigma, bias):
    return GenTailFn.apply(x, weight_orig, inv_sigma, bias)


def tail_tensor_core_ok(x, weight_orig):
    """The generator tail runs on the tensor cores when the halo kernel takes the shape (Cin % 64 == 0, power-of-two
    plane >= 32 x 16, 4 output channels); otherwise the CUDA-core kernels of GenTailFn are used."""
    n, h, w, c = x.shape
    pow2 = (h & (h - 1)) == 0 and (w & (w - 1)) == 0
    return c % 64 == 0 and pow2 and h >= 32 and w >= 16 and tuple(weight_orig.shape[2:]) == (3, 3) and weight_orig.shape[0] == 4


class AdaINTailFn(torch.autograd.Function):
    """Final AdaIN + ReLU + tail conv3x3(Cin -> 4) + bias + tanh + rgb*segm composition as ONE autograd node
    (generators/vector_pose_unsupervised_segmentation_noBottleneck.py:80-88,165-181), the conv on the tensor cores:
        forward : in_stats -> adain_relu (bf16 (hi, lo) operand planes) -> tcgen05 bf16x3 conv on the weight zero-padded to
                  32 output channels -> gen_tail_compose (bias, tanh, composition, NCHW images)
        backward: gen_tail_bwd_act (32-channel padded, tf32) -> TF32 data-gradient conv 32 -> Cin, TF32 weight gradient
                  -> AdaIN backward kernels.
    The CUDA-core tail (GenTailFn) needed 349 us forward / 183 us data-gradient at bs 8, 256^2: N = 4 is HBM-bound on
    paper, but 2304 FMAs per pixel from shared-memory weights are LDS-bound in practice."""

    @staticmethod
    def forward(ctx, x, gamma, beta, weight_orig, inv_sigma, bias, eps):
        need_bwd = any(ctx.needs_input_grad)
        if need_bwd:
            mean, rstd, (a_f32, a_split) = K.adain_stats_apply(x, gamma, beta, eps, round_tf32=True, want_f32=True,
                                                               want_split=True)
        else:
            mean, rstd, a_split = K.adain_stats_apply(x, gamma, beta, eps, want_f32=False, want_split=True)
            a_f32 = None
        w32 = torch.zeros((32,) + tuple(weight_orig.shape[1:]), dtype=weight_orig.dtype, device=x.device)
        w32[:4].copy_(weight_orig.detach())
        wp = K.pack_conv_weight(w32, precision=K.BF16X3)
        a32 = K.conv_fwd(a_split, wp, 3, scale=inv_sigma)
        rgbs, segm, t = K.gen_tail_compose(a32, bias.detach())
        ctx.save_for_backward(x, mean, rstd, gamma, beta, weight_orig, inv_sigma, a_f32, t, w32)
        return rgbs, segm

    @staticmethod
    def backward(ctx, d_rgbs, d_segm):
        x, mean, rstd, gamma, beta, weight_orig, inv_sigma, a_f32, t, w32 = ctx.saved_tensors
        d_rgbs = d_rgbs.contiguous() if d_rgbs is not None else None
        d_segm = d_segm.contiguous() if d_segm is not None else None
        need_x, need_g, need_bt, need_w, need_s, need_b = ctx.needs_input_grad[:6]
        da = K.gen_tail_bwd_act(t, d_rgbs, d_segm, stride=32)
        dx = dgm = dbt = dw = ds = db = None
        if need_x or need_g or need_bt:
            wpt = K.pack_conv_weight(w32, transpose=True)             # (Cin, 9, 32): data-gradient layout, tf32
            d_a = K.conv_fwd(da, wpt, 3, scale=inv_sigma)
            # dx is the last decoder block's output gradient: operand of its tf32 gradient MMAs -> stored rounded
            dx, dgm, dbt = K.adain_relu_bwd(x, mean, rstd, gamma, beta, d_a, round_tf32=True)
        if need_w or need_s:
            g = K.conv_wgrad(a_f32, da, 3)[:4].contiguous()
            if need_s:
                ds = _dot(g, weight_orig).reshape(inv_sigma.shape)
            if need_w:
                dw = g * inv_sigma
        if need_b:
            db = K.bias_grad(da)[:4].contiguous()
        return dx, dgm, dbt, dw, ds, db, None


def adain_tail(x, gamma, beta, weight_orig, inv_sigma, bias, eps=1e-4):
    return AdaINTailFn.apply(x, gamma, beta, weight_orig, inv_sigma, bias, eps)


class ReluRoundFn(torch.autograd.Function):
    """tf32(relu(x)) — the discriminator blocks' leading ReLU(inplace) (blocks.py:73 with norm_layer='none')."""

    @staticmethod
    def forward(ctx, x):
        y = K.relu_round(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return K.relu_bwd(y, dy.contiguous())


def relu_round(x):
    return ReluRoundFn.apply(x)


class DiscBlocksFn(torch.autograd.Function):
    """The discriminator's chain of ResBlock(norm_layer='none') as ONE autograd node (discriminators/no_landmarks.py:96-100
    over generators/common/blocks.py:47-111): outputs = the 6 in-place-ReLU'd block inputs (the stored features) + the last
    block's output.  Forward = the kernels of blocks.PlainResBlock; the hand-scheduled backward removes what autograd
    cannot fuse across nodes:
      * a block input feeds the main branch, the skip branch and the feature-matching loss — the three gradients are merged
        by kernels (skip gradient = residual operand of the first conv's data-gradient epilogue, read at half resolution
        behind a pooled 1x1 skip; feature-matching gradient = addend of the ReLU-backward pass) instead of two
        full-tensor `add` launches per block;
      * ReLU backward, the bias gradients of the convolutions that produced the tensor (column sums at the LOW resolution
        for pooled outputs) and the 0.25-scaled copy the pooled skip needs are one pass (b200lp_relu_bwd_fused).
    `specs`: per block dict(down=bool, c0=(w, s, b, cache, sn), c1=(...), sk=(...) or None); weights / biases are ALSO
    passed in `params` (block order: c0.w, c0.b, c1.w, c1.b[, sk.w, sk.b]) so that autograd routes their gradients when no
    gradient sink takes them."""

    @staticmethod
    def forward(ctx, x0, specs, *params):
        ctx.set_materialize_grads(False)
        out = x0.contiguous()
        feats, keep = [], []
        for sp in specs:
            w0, s0, b0, c0, _ = sp["c0"]
            w1, s1, b1, c1, _ = sp["c1"]
            r = K.relu_round(out)
            feats.append(r)
            h = K.conv_fwd(r, _packed(w0, c0, False), 3, bias=b0, relu=True, round_tf32=True, scale=s0)
            rs = None
            if sp["sk"] is not None:
                ws, ss, bs, cs, _ = sp["sk"]
                rs = K.avgpool2(r, None, round_tf32=True) if sp["down"] else r
                s = K.conv_fwd(rs, _packed(ws, cs, False), 1, bias=bs, scale=ss)
            else:
                s = r
            if sp["down"]:
                out = K.avgpool2(K.conv_fwd(h, _packed(w1, c1, False), 3, bias=b1, scale=s1), s)
            else:
                out = K.conv_fwd(h, _packed(w1, c1, False), 3, bias=b1, residual=s, residual_mode=1, scale=s1)
            keep.append((h, rs))
        ctx.specs = specs
        ctx.n_params = len(params)
        ctx.out_shape = tuple(out.shape)
        flat = list(feats) + [h for h, _ in keep] + [rs for _, rs in keep if rs is not None]
        ctx.has_rs = [rs is not None for _, rs in keep]
        ctx.save_for_backward(*flat)
        return (*feats, out)

    @staticmethod
    def backward(ctx, *grads):
        specs = ctx.specs
        n = len(specs)
        saved = ctx.saved_tensors
        feats = saved[:n]
        rs_it = iter(saved[2 * n:])
        keep = [(saved[n + i], next(rs_it) if ctx.has_rs[i] else None) for i in range(n)]
        d_feats, d_o = grads[:n], grads[n]
        need_x = ctx.needs_input_grad[0]
        pgrads = [None] * ctx.n_params
        pidx = []                       # first parameter slot of every block
        k = 0
        for sp in specs:
            pidx.append(k)
            k += 6 if sp["sk"] is not None else 4
        wants = ctx.needs_input_grad[2:]
        # gradients that only feed tf32 MMAs are stored rounded to nearest (the MMA would truncate: a 2^-12 relative bias
        # per layer that compounds along the chain); B200LP_DISC_NODE_TRUNCATE=1 (A/B test against the per-layer nodes) keeps
        # the unrounded values
        rnd = _round_gradients()

        def bias_into(b, slot, dy):
            """bias gradient by a separate pass (the chain's end, or no sink): column sums of dy."""
            if b is None or not wants[slot]:
                return
            sink = _sink(b)
            if sink is not None:
                K.bias_grad(dy, acc_into=sink)
            else:
                g = K.bias_grad(dy)
                pgrads[slot] = g if pgrads[slot] is None else pgrads[slot] + g

        def wgrad(conv, slot, ksize, x, dy):
            w, s, _, cache, sn = conv
            if not wants[slot]:
                return
            _, dw, _ = _conv_backward(ksize, x, w, s, dy, False, True, False, cache, sn)
            pgrads[slot] = dw              # None when a sink took it

        if d_o is None:                    # only features were used downstream
            d_o = torch.zeros(ctx.out_shape, dtype=torch.float32, device=feats[0].device)
        d_o = d_o.contiguous()
        q = None                           # 0.25 * d_o when the current block is pooled and has a skip conv
        fused_bias = False                 # True when the previous pass already accumulated this block's c1 / sk bias sums
        dx0 = None
        for i in reversed(range(n)):
            sp = specs[i]
            h, rs = keep[i]
            r = feats[i]
            p0 = pidx[i]
            w0, s0, b0, c0, _ = sp["c0"]
            w1, s1, b1, c1, _ = sp["c1"]
            if not fused_bias:
                bias_into(b1, p0 + 3, d_o)
                if sp["sk"] is not None:
                    bias_into(sp["sk"][2], p0 + 5, d_o)
            # second conv (+ avg-pool): data gradient, weight gradient
            dh2 = K.avgpool2_bwd(d_o) if sp["down"] else d_o
            dh = K.conv_fwd(dh2, _packed(w1, c1, True), 3, scale=s1)
            wgrad(sp["c1"], p0 + 2, 3, h, dh2)
            del dh2
            # ReLU between the convs + bias gradient of the first conv, one pass
            b0_sink = _sink(b0) if (b0 is not None and wants[p0 + 1]) else None
            dh_m = K.relu_bwd_fused(h, dh, bias_a=b0_sink, round_tf32=rnd)
            del dh
            if b0_sink is None:
                bias_into(b0, p0 + 1, dh_m)
            # skip branch: its gradient w.r.t. the block input becomes the residual operand of the first conv's data gradient
            if sp["sk"] is not None:
                ws, ss, bs, cs, _ = sp["sk"]
                if sp["down"] and q is not None:
                    d_skip, mode = K.conv_fwd(q, _packed(ws, cs, True), 1, scale=ss), 2          # stays at half resolution
                else:
                    d_skip, mode = K.conv_fwd(d_o, _packed(ws, cs, True), 1, scale=ss), 1
                    if sp["down"]:
                        d_skip = K.avgpool2_bwd(d_skip)
                wgrad(sp["sk"], p0 + 4, 1, rs, d_o)
            else:
                d_skip, mode = d_o, 1
            last = i == 0
            if need_x or not last:
                dr = K.conv_fwd(dh_m, _packed(w0, c0, True), 3, scale=s0, residual=d_skip, residual_mode=mode)
            wgrad(sp["c0"], p0, 3, r, dh_m)
            del dh_m, d_skip
            if last:
                if need_x:
                    dx0 = K.relu_bwd_fused(r, dr, add=d_feats[0].contiguous() if d_feats[0] is not None else None)
                break
            # the block input's in-place ReLU + feature-matching gradient + bias sums / quarter copy for the block before
            pv = specs[i - 1]
            pp = pidx[i - 1]
            ba = _sink(pv["c1"][2]) if (pv["c1"][2] is not None and wants[pp + 3]) else None
            bb = _sink(pv["sk"][2]) if (pv["sk"] is not None and pv["sk"][2] is not None and wants[pp + 5]) else None
            want_q = pv["down"] and pv["sk"] is not None
            add = d_feats[i].contiguous() if d_feats[i] is not None else None
            res = K.relu_bwd_fused(r, dr, add=add, want_quarter=want_q, bias_a=ba, bias_b=bb, round_tf32=rnd)
            d_o, q = res if want_q else (res, None)
            del dr
            fused_bias = True
            # biases the fused pass could not take (no sink): separate column sums
            if ba is None:
                bias_into(pv["c1"][2], pp + 3, d_o)
            if pv["sk"] is not None and bb is None:
                bias_into(pv["sk"][2], pp + 5, d_o)
        return (dx0, None, *pgrads)


def disc_blocks(x0, specs):
    params = []
    for sp in specs:
        for key in ("c0", "c1", "sk"):
            if sp[key] is not None:
                params += [sp[key][0], sp[key][2]]
    outs = DiscBlocksFn.apply(x0, specs, *params)
    return list(outs[:-1]), outs[-1]


class AvgPool2Fn(torch.autograd.Function):
    """avg_pool2d(x, 2) (+ addend) — nn.AvgPool2d(2) at blocks.py:89-90,101-102, discriminator :62,66."""

    @staticmethod
    def forward(ctx, x, addend, round_out):
        ctx.has_add = addend is not None
        return K.avgpool2(x, addend, round_tf32=round_out)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        dx = K.avgpool2_bwd(dy) if ctx.needs_input_grad[0] else None
        return dx, (dy if ctx.has_add and ctx.needs_input_grad[1] else None), None


def avgpool2(x, addend=None, round_out=False):
    return AvgPool2Fn.apply(x, addend, round_out)


class ConvC3Fn(torch.autograd.Function):
    """3x3 conv on a 3-channel NCHW image -> NHWC features (discriminator down_block.0, skip.0 via a centre-tap
    embedding; VGG features.0 with the caffe input normalisation folded in)."""

    @staticmethod
    def forward(ctx, x_nchw, weight_orig, inv_sigma, bias, pre_scale, pre_shift, relu, round_out):
        y = K.conv3x3_c3_fwd(x_nchw, weight_orig, inv_sigma, bias, pre_scale, pre_shift, relu=relu, round_tf32=round_out)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.bias_ref = bias.detach() if bias is not None else None      # alias only: identifies the gradient sink
        ctx.save_for_backward(x_nchw, weight_orig, inv_sigma, pre_scale, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight_orig, inv_sigma, pre_scale, y = ctx.saved_tensors
        dy = dy.contiguous()
        need_x, need_w, need_s, need_b = ctx.needs_input_grad[:4]
        bias_done = False
        if ctx.relu:
            bsink = _sink(ctx.bias_ref) if (ctx.has_bias and need_b) else None
            if bsink is not None and y.shape[-1] % 4 == 0 and y.shape[-1] <= 1024:
                dy = K.relu_bwd_fused(y, dy, bias_a=bsink, round_tf32=_round_gradients())     # mask + bias sums + tf32 rounding, one pass
                bias_done = True
            else:
                dy = K.relu_bwd(y, dy)
        dx = dw = ds = db = None
        if need_x:
            dx = K.conv3x3_c3_dgrad_tc(dy, K.c3_transposed_weight(weight_orig), inv_sigma, pre_scale)
        if need_w or need_s:
            assert pre_scale is None, "weight gradient with input pre-affine is not needed on this path"
            g = K.conv3x3_c3_wgrad_tc(x, dy)
            if inv_sigma is not None:
                if need_s:
                    ds = _dot(g, weight_orig).reshape(inv_sigma.shape)
                if need_w:
                    dw = g * inv_sigma
            else:
                dw = g
        if ctx.has_bias and need_b and not bias_done:
            db = K.bias_grad(dy)
        return dx, dw, ds, db, None, None, None, None


def conv_c3(x_nchw, weight_orig, inv_sigma=None, bias=None, pre_scale=None, pre_shift=None, relu=False,
            round_out=False):
    return ConvC3Fn.apply(x_nchw, weight_orig, inv_sigma, bias, pre_scale, pre_shift, relu, round_out)


class L1MeanFn(torch.autograd.Function):
    """F.l1_loss(a, b.detach()) (mean reduction) — criterions/featmat.py:18-20, perceptual_loss.py:108."""

    @staticmethod
    def forward(ctx, a, b):
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
        K.l1_sum(a, b, out, 1.0 / a.numel())
        ctx.save_for_backward(a, b)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        gs = g.reshape(1).contiguous().float()
        return K.l1_bwd(a, b, gs, 1.0 / a.numel()), None


def l1_mean(a, b):
    return L1MeanFn.apply(a.contiguous(), b.detach().contiguous())


class L1MeanSumFn(torch.autograd.Function):
    """scale * sum_i mean|a_i - b_i| over a list of tensor pairs as ONE node with ONE accumulator (criterions/featmat.py:18-20
    sums F.l1_loss over the 7 discriminator feature maps): 7 reduction kernels forward, 7 sign kernels backward, and none of
    the per-pair zero-fills / scalar adds / multiplies of the per-pair form."""

    @staticmethod
    def forward(ctx, scale, n, *tensors):
        a, b = tensors[:n], tensors[n:]
        out = torch.zeros(1, dtype=torch.float32, device=a[0].device)
        for x, y in zip(a, b):
            K.l1_sum(x, y, out, scale / x.numel())
        ctx.scale, ctx.n = scale, n
        ctx.save_for_backward(*tensors)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        n = ctx.n
        a, b = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        gs = g.reshape(1).contiguous().float()
        grads = [K.l1_bwd(x, y, gs, ctx.scale / x.numel()) if ctx.needs_input_grad[2 + i] else None
                 for i, (x, y) in enumerate(zip(a, b))]
        return (None, None, *grads, *([None] * n))


def l1_mean_sum(pairs, scale):
    """scale * sum over pairs of mean|a - b.detach()|."""
    a = [x.contiguous() for x, _ in pairs]
    b = [y.detach().contiguous() for _, y in pairs]
    return L1MeanSumFn.apply(float(scale), len(a), *a, *b)


class NchwToNhwcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.nchw_to_nhwc(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return K.nhwc_to_nchw(dy.contiguous())


class NhwcToNchwFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return K.nhwc_to_nchw(x.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return K.nchw_to_nhwc(dy.contiguous())


def nchw_to_nhwc(x):
    return NchwToNhwcFn.apply(x)


def nhwc_to_nchw(x):
    return NhwcToNchwFn.apply(x)


# ----------------------------------------------------------------------------------------------------------------
# VGG perceptual loss as ONE autograd node: frozen weights (packed once), fake and real pushed side by side, L1
# reduced by kernel after every ReLU; backward = data-gradient chain only (criterions/common/perceptual_loss.py:91-110)
# ----------------------------------------------------------------------------------------------------------------
class VggPerceptualFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, fake_nchw, real_nchw, packed, weight):
        """packed: dict from VggFeatures.pack(): plan [(kind, idx)], first conv OIHW + bias, packed fwd/bwd weights,
        biases, pre_scale/pre_shift of the fused input normalisation."""
        plan = packed["plan"]
        dev = fake_nchw.device
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        saved = []          # per ReLU tap: (backward code of l1_sum_code, feature shape, pooled-after flag)
        need_bwd = ctx.needs_input_grad[0]
        # fake and real go through the frozen network as ONE batch of 2B images (first half fake, second half real): half
        # the launches, twice the tiles per launch (the 16 x 16 layers no longer need split-K); a tap compares the halves
        nb = fake_nchw.shape[0]
        both = torch.cat([fake_nchw, real_nchw], dim=0).contiguous()
        f = None
        skip_pool = False
        for pos, (kind, idx) in enumerate(plan):
            if kind == "conv0":   # conv + relu fused (every VGG conv is followed by a ReLU tap)
                f = K.conv3x3_c3_fwd(both, packed["w0"], None, packed["b0"], packed["pre_scale"], packed["pre_shift"],
                                     relu=True, round_tf32=True)
            elif kind == "conv":
                f = K.conv_fwd(f, packed["wp"][idx], 3, bias=packed["bias"][idx], relu=True, round_tf32=True)
            elif kind == "pool":
                if skip_pool:     # already produced by the fused tap before it
                    skip_pool = False
                    continue
                f = K.avgpool2(f, None, round_tf32=True)
                continue
            a, b = f[:nb], f[nb:]
            pool_next = (need_bwd and pos + 1 < len(plan) and plan[pos + 1][0] == "pool"
                         and a.shape[1] % 2 == 0 and a.shape[2] % 2 == 0)
            if pool_next:   # tap + the pool behind it in one pass over both feature maps
                shape = tuple(a.shape)
                pooled = torch.empty((2 * nb, shape[1] // 2, shape[2] // 2, shape[3]), dtype=torch.float32, device=dev)
                code, _, _ = K.l1_sum_code_pool(a, b, loss, weight / a.numel(), ap=pooled[:nb], bp=pooled[nb:])
                f = pooled
                saved.append((code, shape, True))
                skip_pool = True
            elif need_bwd:  # the L1 term and, in the same pass, the 2-bit (ReLU mask, sign) code the backward tap needs
                saved.append((K.l1_sum_code(a, b, loss, weight / a.numel()), tuple(a.shape), False))
            else:
                K.l1_sum(a, b, loss, weight / a.numel())
        ctx.packed = packed
        ctx.weight = weight
        ctx.saved_feats = saved
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        packed, saved = ctx.packed, ctx.saved_feats
        plan = packed["plan"]
        gs = g.reshape(1).contiguous().float()
        d = None                     # gradient w.r.t. the current activation (post-ReLU feature), NHWC
        tap = len(saved) - 1
        dx_img = None
        pending_pool = False         # a pool entry was passed whose un-pooling the tap before it will do
        for kind, idx in reversed(plan):
            if kind == "pool":
                if tap >= 0 and saved[tap][2]:
                    pending_pool = True
                else:
                    d = K.avgpool2_bwd(d)
                continue
            code, shape, pooled = saved[tap]
            tap -= 1
            numel = 1
            for v in shape:
                numel *= v
            # d(loss)/d(a) gets the L1 term of this tap, then passes the ReLU mask (one pass over the code + d)
            if pooled and pending_pool and d is not None:
                d = K.l1_code_bwd_unpool(code, shape, gs, ctx.weight / numel, d)
            else:
                if pending_pool and d is not None:
                    d = K.avgpool2_bwd(d)
                d = K.l1_code_bwd(code, shape, gs, ctx.weight / numel, d_in=d)
            pending_pool = False
            if kind == "conv":
                d = K.conv_fwd(d, packed["wpt"][idx], 3)
            else:
                dx_img = K.conv3x3_c3_dgrad_tc(d, packed["w0t"], None, packed["pre_scale"])
        return dx_img, None, None, None


def vgg_perceptual(fake_nchw, real_nchw, packed, weight):
    return VggPerceptualFn.apply(fake_nchw, real_nchw.detach(), packed, float(weight))


# ----------------------------------------------------------------------------------------------------------------
# Small dense layers and scalar losses (csrc/losses.cu, sgemm of csrc/encoder.cu): the last torch arithmetic of the step
# ----------------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = inv_sigma * (x W^T) + b — nn.Linear under spectral_norm without materialising W / sigma
    (generators/vector_pose_unsupervised_segmentation_noBottleneck.py:97-101: the 768 -> 768 -> 13056 projector)."""

    @staticmethod
    def forward(ctx, x, weight, inv_sigma, bias):
        x = x.contiguous()
        y = K.sgemm(x, weight.detach(), trans_b=True, alpha=inv_sigma, bias=bias.detach() if bias is not None else None)
        ctx.save_for_backward(x, weight, inv_sigma)
        ctx.bias_ref = bias.detach() if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, inv_sigma = ctx.saved_tensors
        dy = dy.contiguous()
        need_x, need_w, need_s, need_b = ctx.needs_input_grad
        dx = dw = ds = db = None
        if need_x:
            dx = K.sgemm(dy, weight.detach(), alpha=inv_sigma)
        if need_w or need_s:
            sink = _sink(weight) if need_w else None
            if sink is not None and not need_s:
                K.sgemm(dy, x, trans_a=True, acc_into=sink, alpha=inv_sigma)
            else:
                g = K.sgemm(dy, x, trans_a=True)                    # dL/d(W / sigma) before the scale
                if need_s:
                    ds = _dot(g, weight.detach()).reshape(inv_sigma.shape)
                if need_w:
                    if sink is not None:
                        sink.add_(g * inv_sigma)
                    else:
                        dw = g * inv_sigma
        if need_b and ctx.bias_ref is not None:
            bsink = _sink(ctx.bias_ref)
            if bsink is not None:
                K.bias_grad(dy, acc_into=bsink)
            else:
                db = K.bias_grad(dy)
        return dx, dw, ds, db


def linear(x, weight, inv_sigma, bias):
    return LinearFn.apply(x, weight, inv_sigma, bias)


class DiscHeadFn(torch.autograd.Function):
    """score = inv_sigma * <o, w> + b + <o, embed>, o = spatial sum of relu(feat) — the discriminator's projection head
    (discriminators/no_landmarks.py:101-105) as one kernel forward, two backward."""

    @staticmethod
    def forward(ctx, feat, embed, weight, inv_sigma, bias):
        feat = feat.contiguous()
        emb = embed.contiguous() if embed is not None else None
        score, o = K.disc_head_fwd(feat, emb, weight.detach().reshape(-1), inv_sigma, bias.detach())
        ctx.save_for_backward(feat, emb, weight, inv_sigma, o)
        ctx.bias_ref = bias.detach()
        return score

    @staticmethod
    def backward(ctx, g):
        feat, emb, weight, inv_sigma, o = ctx.saved_tensors
        need_f, need_e, need_w, need_s, need_b = ctx.needs_input_grad
        wsink = _sink(weight) if need_w else None
        bsink = _sink(ctx.bias_ref) if need_b else None
        in_place = wsink is not None and bsink is not None
        d_feat, d_emb, dw, ds, db = K.disc_head_bwd(
            feat, emb, weight.detach().reshape(-1), inv_sigma, o, g.contiguous(), need_feat=need_f or need_e,
            need_embed=need_e, need_params=need_w or need_s or need_b,
            dw_acc=wsink.reshape(-1) if in_place else None, db_acc=bsink if in_place else None)
        if in_place:
            dw = db = None
        else:
            if dw is not None and wsink is not None:
                wsink.add_(dw.view_as(wsink)); dw = None
            if db is not None and bsink is not None:
                bsink.add_(db.view_as(bsink)); db = None
        return (d_feat if need_f else None, d_emb if need_e else None,
                dw.view_as(weight) if (dw is not None and need_w) else None,
                ds.reshape(inv_sigma.shape) if (ds is not None and need_s) else None,
                db if need_b else None)


def disc_head(feat, embed, weight, inv_sigma, bias):
    return DiscHeadFn.apply(feat, embed, weight, inv_sigma, bias)


class DiceFn(torch.autograd.Function):
    """-log(2 sum(f*r) / (sum f^2 + sum r^2)) * w with f (B,1,H,W) broadcast over r's channels (criterions/dice.py:30-34)."""

    @staticmethod
    def forward(ctx, fake_segm, real_segm, weight):
        f, r = fake_segm.contiguous(), real_segm.contiguous()
        loss, sums = K.dice_fwd(f, r, weight)
        ctx.save_for_backward(f, r, sums)
        ctx.weight = weight
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        f, r, sums = ctx.saved_tensors
        return K.dice_bwd(f, r, sums, g.reshape(1).contiguous().float(), ctx.weight), None, None


def dice_loss(fake_segm, real_segm, weight):
    return DiceFn.apply(fake_segm, real_segm.detach(), float(weight))


class AdversarialGFn(torch.autograd.Function):
    """loss_G = -mean(fake_score_G) (criterions/adversarial.py:44-45, `gan`).  Separate nodes for the two losses:
    loss_D.backward() must not reach the generator's graph through a shared node."""

    @staticmethod
    def forward(ctx, fake_g):
        fg = fake_g.contiguous()
        ctx.n = fg.numel()
        ctx.save_for_backward(fg)
        return K.adversarial_fwd(fg, fg, fg, 0)[0]

    @staticmethod
    def backward(ctx, g):
        (fg,) = ctx.saved_tensors
        dg, _, _ = K.adversarial_bwd(fg, fg, g.reshape(1).contiguous().float(), None, need_g=True, need_d=False)
        return dg


class AdversarialDFn(torch.autograd.Function):
    """loss_D = mean(relu(1 - real)) + mean(relu(1 + fake_D)) (criterions/adversarial.py:42, hinge)."""

    @staticmethod
    def forward(ctx, fake_d, real):
        fd, rl = fake_d.contiguous(), real.contiguous()
        ctx.save_for_backward(fd, rl)
        return K.adversarial_fwd(fd, fd, rl, 0)[1]

    @staticmethod
    def backward(ctx, g):
        fd, rl = ctx.saved_tensors
        _, dd, dr = K.adversarial_bwd(fd, rl, None, g.reshape(1).contiguous().float(), need_g=False, need_d=True)
        return dd, dr


def adversarial_losses(fake_g, fake_d, real):
    return AdversarialGFn.apply(fake_g), AdversarialDFn.apply(fake_d, real)


class CropFn(torch.autograd.Function):
    """Box crop + bilinear resize (criterions/idt_embed.py:62-83, torch affine_grid + grid_sample(bilinear, reflection))."""

    @staticmethod
    def forward(ctx, images, boxes):
        x = images.contiguous()
        ctx.save_for_backward(boxes)
        ctx.hw = tuple(x.shape[2:])
        return K.crop_bilinear_fwd(x, boxes)

    @staticmethod
    def backward(ctx, dy):
        (boxes,) = ctx.saved_tensors
        return K.crop_bilinear_bwd(dy.contiguous(), boxes, ctx.hw), None


def crop_boxes_inside(boxes_host, h, w, oh, ow):
    """True when every sampling position of the boxes (host list of [t, b, l, r]) lies inside the image, i.e. the
    reflection padding never acts and CropFn's gather-form backward is exact."""
    for t, b, l, r in boxes_host:
        ay, ax = (b - t) / oh, (r - l) / ow
        if ay <= 0 or ax <= 0:
            return False
        y0, y1 = t + 0.5 * ay - 0.5, t + 0.5 * ay - 0.5 + ay * (oh - 1)
        x0, x1 = l + 0.5 * ax - 0.5, l + 0.5 * ax - 0.5 + ax * (ow - 1)
        if y0 < 0 or x0 < 0 or y1 > h - 1 or x1 > w - 1:
            return False
    return True


def crop_bilinear(images, boxes):
    return CropFn.apply(images, boxes)
