"""Pose encoder: a torchvision MobileNetV2 module run forward AND backward as a schedule of libb200lp kernels
(csrc/mobilenet.cu, csrc/mobilenet_bwd.cu, BatchNorm backward of csrc/encoder.cu).

The module keeps owning every parameter and buffer (checkpoint keys `pose_encoder.*` are torchvision's); this file only
sequences kernels over them: conv -> (batch statistics) -> bn_finalize, with every BatchNorm + ReLU6 applied by the
CONSUMER kernel on load.  Same results as `net(x)` (reference: embedders/unsupervised_pose_separate_embResNeXt_
segmentation.py:56-58) including the train-mode running-statistics updates.  Backward (meta-training, where the pose
encoder is in optimizer_G): BatchNorm backward with the ReLU6 mask recomputed from the raw conv outputs, depthwise
data / weight gradients, 1x1 data gradients as the forward SGEMM on the transposed weight, 1x1 weight gradients as a
row-split dy^T x product with the producer's BatchNorm + ReLU6 applied on load; parameter gradients go straight into the
runner's gradient bucket when `b200lp.ops.direct_grads` is active.  One autograd node (`PoseFn`) spans the network.
The classifier's Dropout(0.2) is RNG-dependent and stays a torch op on the (N, 1280) pooled features.
"""
import torch

from b200lp import kernels as K
from b200lp import ops


def _conv_bn(seq):
    """(conv, bn) of a torchvision Conv2dNormActivation / [conv, bn] pair."""
    return seq[0], seq[1]


def supported(net):
    """True if `net` has the layout this schedule walks (torchvision.models.MobileNetV2, width 1.0)."""
    try:
        f = net.features
        conv0, bn0 = _conv_bn(f[0])
        ok = tuple(conv0.weight.shape) == (32, 3, 3, 3) and conv0.stride == (2, 2) and conv0.bias is None
        ok = ok and isinstance(bn0, torch.nn.BatchNorm2d) and isinstance(net.classifier[1], torch.nn.Linear)
        for blk in list(f)[1:-1]:
            layers = list(blk.conv)
            ok = ok and len(layers) in (3, 4) and hasattr(blk, 'use_res_connect')
        for m in net.modules():      # the finalize kernel implements the exponential running average only
            if isinstance(m, torch.nn.BatchNorm2d):
                ok = ok and m.momentum is not None and m.affine
        return bool(ok)
    except Exception:
        return False


class _BN:
    """What a BatchNorm layer normalised with in this forward pass (inputs of the on-load apply and of bn_bwd)."""
    __slots__ = ("mod", "scale", "shift", "mean", "rstd", "batch_stats")

    def __init__(self, mod, part, count, want_stats):
        self.mod = mod
        self.batch_stats = mod.training or not mod.track_running_stats
        if want_stats:
            self.scale, self.shift, self.mean, self.rstd = K.bn_finalize(mod, part if self.batch_stats else None, count,
                                                                         self.batch_stats, want_stats=True)
        else:
            self.scale, self.shift = K.bn_finalize(mod, part if self.batch_stats else None, count, self.batch_stats)
            self.mean = self.rstd = None


def _stats_on(bn):
    return bn.training or not bn.track_running_stats


def forward(net, x_nchw, need_bwd=False):
    """x_nchw: (N, 3, H, W) float32 CUDA -> (N, num_classes) [, saved state when need_bwd].  Honours net.training (batch
    statistics + running-stat updates + dropout) vs eval (running statistics)."""
    x = x_nchw.contiguous().float()
    f = net.features
    saved = {"x": x, "blocks": []} if need_bwd else None

    def pw(inp, inp_bn, relu6, conv, bn):
        nn_, h, w, c = inp.shape
        wt = conv.weight.detach().reshape(conv.out_channels, conv.in_channels)
        sc, sh = (inp_bn.scale, inp_bn.shift) if inp_bn is not None else (None, None)
        if _stats_on(bn):
            y, part = K.pw_conv(inp.reshape(-1, c), wt, sc, sh, relu6, want_stats=True)
        else:
            y, part = K.pw_conv(inp.reshape(-1, c), wt, sc, sh, relu6), None
        return y.reshape(nn_, h, w, conv.out_channels), _BN(bn, part, y.shape[0], need_bwd)

    # stem: conv3x3 s2 -> lazy (raw, bn): consumers apply BN + ReLU6 on load
    conv, bn = _conv_bn(f[0])
    if _stats_on(bn):
        raw, part = K.mbv2_stem(x, conv.weight.detach(), want_stats=True)
    else:
        raw, part = K.mbv2_stem(x, conv.weight.detach()), None
    stem_bn = _BN(bn, part, raw.numel() // raw.shape[-1], need_bwd)
    cur, cur_bn = raw, stem_bn             # lazy activation relu6(bn(cur)); materialised when cur_bn is None
    if need_bwd:
        saved.update(stem_raw=raw, stem_bn=stem_bn, stem_conv=conv)

    for blk in list(f)[1:-1]:
        layers = list(blk.conv)
        block_in, block_in_bn = cur, cur_bn
        rec = dict(blk=blk, inp=cur, inp_bn=cur_bn)
        if len(layers) == 4:                            # expand 1x1 + BN + ReLU6
            conv, bn = _conv_bn(layers[0])
            e, e_bn = pw(cur, cur_bn, cur_bn is not None, conv, bn)
            rec.update(e_conv=conv)
            dw_seq = layers[1]
        else:                                           # t = 1 block: depthwise acts on the incoming lazy activation
            assert cur_bn is not None, "depthwise conv needs a BatchNorm+ReLU6 producer"
            e, e_bn = cur, cur_bn
            dw_seq = layers[0]
        conv, bn = _conv_bn(dw_seq)
        assert conv.groups == conv.in_channels == conv.out_channels and conv.kernel_size == (3, 3)
        if _stats_on(bn):
            d, part = K.dw_conv3x3(e, conv.weight.detach(), e_bn.scale, e_bn.shift, conv.stride[0], want_stats=True)
        else:
            d, part = K.dw_conv3x3(e, conv.weight.detach(), e_bn.scale, e_bn.shift, conv.stride[0]), None
        d_bn = _BN(bn, part, d.numel() // d.shape[-1], need_bwd)
        rec.update(e=e, e_bn=e_bn, dw_conv=conv, d=d, d_bn=d_bn)
        conv, bn = layers[-2], layers[-1]               # linear 1x1 + BN
        p, p_bn = pw(d, d_bn, True, conv, bn)
        res = None
        if blk.use_res_connect:
            assert block_in_bn is None
            res = block_in
        cur, cur_bn = K.bn_apply(p, p_bn.scale, p_bn.shift, residual=res), None
        if need_bwd:
            rec.update(p_conv=conv, p=p, p_bn=p_bn, residual=res is not None)
            saved["blocks"].append(rec)

    conv, bn = _conv_bn(f[-1])                          # 1x1 -> 1280, BN, ReLU6, global average pool
    last, l_bn = pw(cur, cur_bn, cur_bn is not None, conv, bn)
    pooled = K.bn_relu6_avgpool(last, l_bn.scale, l_bn.shift)
    drop, lin = net.classifier[0], net.classifier[1]
    dropped = pooled
    leaf = None
    if net.training and isinstance(drop, torch.nn.Dropout) and drop.p > 0:
        if need_bwd:        # RNG-dependent, (N, 1280): a torch op; its tiny backward goes through torch autograd
            leaf = pooled.detach().requires_grad_(True)
            with torch.enable_grad():
                dropped = torch.nn.functional.dropout(leaf, drop.p, True)
        else:
            dropped = torch.nn.functional.dropout(pooled, drop.p, True)
    out = K.pw_conv(dropped.detach(), lin.weight.detach(), bias=lin.bias.detach() if lin.bias is not None else None)
    if need_bwd:
        saved.update(last_conv=conv, last_in=cur, last=last, l_bn=l_bn, leaf=leaf, dropped=dropped, lin=lin)
        return out, saved
    return out


class _Grads:
    """Parameter gradients of one backward pass: into the active gradient sinks or collected for autograd."""

    def __init__(self, params):
        self.index = {id(p): i for i, p in enumerate(params)}
        self.out = [None] * len(params)

    def give(self, p, g):
        s = ops._sink(p)
        if s is not None:
            s.add_(g.view_as(s))
        else:
            i = self.index[id(p)]
            self.out[i] = g.view_as(p) if self.out[i] is None else self.out[i] + g.view_as(p)


def _bn_backward(grads, st, dy, x_raw, mask_mode):
    """BatchNorm (+ReLU6 when mask_mode 3) backward; gamma / beta gradients to the sinks or autograd."""
    bn = st.mod
    need = bn.weight.requires_grad
    sg, sb = (ops._sink(bn.weight), ops._sink(bn.bias)) if need else (None, None)
    if sg is not None and sb is not None:
        dx, _, _, _ = K.bn_bwd(dy, x_raw, st.mean, st.rstd, bn.weight.detach(), st.scale, st.shift, mask_mode=mask_mode,
                               dgamma=sg, dbeta=sb, accumulate=True, batch_stats=st.batch_stats)
    else:
        dx, dg, db, _ = K.bn_bwd(dy, x_raw, st.mean, st.rstd, bn.weight.detach(), st.scale, st.shift, mask_mode=mask_mode,
                                 batch_stats=st.batch_stats)
        if need:
            grads.give(bn.weight, dg)
            grads.give(bn.bias, db)
    return dx


def _pw_backward(grads, conv, dy, inp, inp_bn, relu6, need_dx=True):
    """1x1 conv backward: weight gradient dy^T f(inp) (f = producer BN [+ReLU6] on load) and data gradient dy W."""
    cout, cin = conv.out_channels, conv.in_channels
    dy2 = dy.reshape(-1, cout)
    if conv.weight.requires_grad:
        sc, sh = (inp_bn.scale, inp_bn.shift) if inp_bn is not None else (None, None)
        sink = ops._sink(conv.weight)
        g = K.pw_wgrad(dy2, inp.reshape(-1, cin), sc, sh, relu6, acc_into=sink.view(cout, cin) if sink is not None else None)
        if sink is None:
            grads.give(conv.weight, g)
    if not need_dx:
        return None
    wt_t = K.transpose2d(conv.weight.detach().reshape(cout, cin))            # (cin, cout): dx = dy @ W as x @ (W^T)^T
    return K.pw_conv(dy2, wt_t).reshape(tuple(dy.shape[:-1]) + (cin,))


def backward(net, saved, d_out, params):
    """d_out (N, num_classes) -> list of parameter gradients aligned with `params` (None where a sink took it)."""
    grads = _Grads(params)
    lin = saved["lin"]
    d_out = d_out.contiguous().float()
    dropped = saved["dropped"].detach()
    if lin.weight.requires_grad:
        sink = ops._sink(lin.weight)
        g = K.pw_wgrad(d_out, dropped, acc_into=sink)
        if sink is None:
            grads.give(lin.weight, g)
    if lin.bias is not None and lin.bias.requires_grad:
        sink = ops._sink(lin.bias)
        g = K.bias_grad(d_out, acc_into=sink)
        if sink is None:
            grads.give(lin.bias, g)
    d_drop = K.pw_conv(d_out, K.transpose2d(lin.weight.detach()))
    if saved["leaf"] is not None:
        (d_pooled,) = torch.autograd.grad(saved["dropped"], saved["leaf"], d_drop)
    else:
        d_pooled = d_drop
    last = saved["last"]
    d_act = K.avgpool_bwd(d_pooled.contiguous(), (last.shape[1], last.shape[2]))
    d_last = _bn_backward(grads, saved["l_bn"], d_act, last, 3)
    d_cur = _pw_backward(grads, saved["last_conv"], d_last, saved["last_in"], None, False)

    blocks = saved["blocks"]
    for bi in range(len(blocks) - 1, -1, -1):
        rec = blocks[bi]
        # block output = bn_p(p) (+ block input): linear BatchNorm backward, the residual passes d_cur through
        dp = _bn_backward(grads, rec["p_bn"], d_cur, rec["p"], 0)
        d_dact = _pw_backward(grads, rec["p_conv"], dp, rec["d"], rec["d_bn"], True)
        dd = _bn_backward(grads, rec["d_bn"], d_dact, rec["d"], 3)
        conv = rec["dw_conv"]
        e, e_bn = rec["e"], rec["e_bn"]
        d_eact = K.dw_dgrad(dd, conv.weight.detach(), (e.shape[1], e.shape[2]), conv.stride[0])
        if conv.weight.requires_grad:
            sink = ops._sink(conv.weight)
            g = K.dw_wgrad(e, dd, e_bn.scale, e_bn.shift, conv.stride[0], acc_into=sink)
            if sink is None:
                grads.give(conv.weight, g)
        de = _bn_backward(grads, e_bn, d_eact, e, 3)          # gradient w.r.t. the raw tensor feeding the depthwise conv
        if "e_conv" in rec:
            inp, inp_bn = rec["inp"], rec["inp_bn"]
            d_in = _pw_backward(grads, rec["e_conv"], de, inp, inp_bn, inp_bn is not None)
            if inp_bn is not None:                            # the block input was the lazy stem output (not in torchvision's
                d_in = _bn_backward(grads, inp_bn, d_in, inp, 3)   # layout: every expand block follows a materialised output)
        else:
            d_in = de              # t = 1 block: e IS the (raw) stem output and e_bn the stem's BatchNorm: already applied
        if rec["residual"]:
            d_in = d_in + d_cur
        d_cur = d_in

    # stem weight gradient: d_cur is the gradient w.r.t. the raw stem conv output
    conv = saved["stem_conv"]
    if conv.weight.requires_grad:
        sink = ops._sink(conv.weight)
        g = K.mbv2_stem_wgrad(saved["x"], d_cur.contiguous(), acc_into=sink)
        if sink is None:
            grads.give(conv.weight, g)
    return grads.out


class PoseFn(torch.autograd.Function):
    """pose embedding = mobilenet_v2(x) as ONE autograd node over the module's parameters."""

    @staticmethod
    def forward(ctx, net, x, *params):
        out, saved = forward(net, x, need_bwd=True)
        ctx.net, ctx.saved, ctx.params = net, saved, params
        return out

    @staticmethod
    def backward(ctx, d_out):
        return (None, None) + tuple(backward(ctx.net, ctx.saved, d_out, ctx.params))


def apply(net, x_nchw):
    """Differentiable (w.r.t. the parameters) forward of `net` on `x_nchw` through the kernel schedule."""
    params = tuple(net.parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return PoseFn.apply(net, x_nchw, *params)
    return forward(net, x_nchw)
