"""Identity encoder: a torchvision ResNeXt50-32x4d module run forward AND backward as a schedule of libb200lp kernels.

Replaces `Embedder.get_identity_embedding`'s network call (reference: embedders/unsupervised_pose_separate_embResNeXt_
segmentation.py:26-27,37-54 — `torchvision.models.resnext50_32x4d(num_classes=512)` on the B*K identity frames, train-mode
BatchNorm over that batch) and its autograd backward.  The module keeps owning every parameter and buffer (checkpoint
keys `identity_encoder.*` are torchvision's); this file only sequences kernels over them:

  * 1x1 convolutions (94 % of the FLOPs) and the 7x7 stem (as a 1x1 GEMM over an im2col patch matrix): tcgen05 implicit
    GEMM — bf16x3 operands forward (three bf16 MMAs per K step on (hi, lo) planes, ~fp32 accuracy: a single-pass TF32
    forward leaves a 2e-2 relative error in the 512-d embedding after 53 layers of batch-normalised convolutions,
    measured against float64), TF32 for the data and weight gradients;
  * grouped 3x3 convolutions: tensor cores on block-diagonal tiles — bf16x3 forward on 64-channel tiles for the 13
    stride-1 layers (`conv_fwd(grouped=cpg)`; operand materialised by `bn_act` and shared with the weight gradient), TF32
    data and weight gradients on 32-channel tiles (`gconv3x3_wgrad_tc`; the three stride-2 layers through the
    zero-stuffed gradient); the three stride-2 forwards and non-power-of-two planes keep the FP32 CUDA-core kernels that
    apply the producer's BatchNorm + ReLU on load and emit their output statistics;
  * BatchNorm: statistics partials -> `bn_finalize` (running-statistics updates like nn.BatchNorm2d) -> one
    materialising pass per GEMM operand; backward = reduce / finalize / apply with the ReLU mask recomputed.

One `torch.autograd.Function` (`ResNeXtFn`) spans the whole network: autograd sees a single node whose inputs are the
module's parameters; parameter gradients are accumulated straight into the runner's gradient bucket when
`b200lp.ops.direct_grads` is active.
"""
import torch

from b200lp import kernels as K
from b200lp import ops


def supported(net):
    """True if `net` has the layout this schedule walks (torchvision ResNet with Bottleneck blocks, groups = 32,
    4..32 channels per group, 7x7/2 stem, 3x3/2 max-pool)."""
    try:
        ok = tuple(net.conv1.weight.shape) == (64, 3, 7, 7) and net.conv1.stride == (2, 2) and net.conv1.bias is None
        ok = ok and net.maxpool.kernel_size == 3 and net.maxpool.stride == 2 and net.maxpool.padding == 1
        ok = ok and isinstance(net.fc, torch.nn.Linear) and net.fc.in_features % 4 == 0 and net.fc.out_features % 4 == 0
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                c2 = blk.conv2
                cpg = c2.in_channels // c2.groups
                ok = ok and c2.kernel_size == (3, 3) and c2.groups == 32 and cpg in (4, 8, 16, 32) and c2.bias is None
                ok = ok and c2.stride[0] in (1, 2) and c2.dilation == (1, 1)
                ok = ok and blk.conv1.kernel_size == (1, 1) and blk.conv3.kernel_size == (1, 1)
                ok = ok and blk.conv1.in_channels % 64 == 0 and blk.conv3.in_channels % 64 == 0
                if blk.downsample is not None:
                    ok = ok and blk.downsample[0].kernel_size == (1, 1) and blk.downsample[0].stride == c2.stride
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                ok = ok and m.momentum is not None and m.affine
        return bool(ok)
    except Exception:
        return False


class _BN:
    """What a BatchNorm layer normalised with in this forward pass."""
    __slots__ = ("mod", "scale", "shift", "mean", "rstd", "batch_stats")

    def __init__(self, mod, part, count):
        self.mod = mod
        self.batch_stats = mod.training or not mod.track_running_stats
        self.scale, self.shift, self.mean, self.rstd = K.bn_finalize(mod, part if self.batch_stats else None, count,
                                                                     self.batch_stats, want_stats=True)


def _stats_needed(bn):
    return bn.training or not bn.track_running_stats


def _packed(conv, transpose, precision):
    cache = conv.__dict__.get("_b200lp_pack")
    if cache is None:
        cache = conv.__dict__["_b200lp_pack"] = ops.PackCache()
    return cache.get(conv.weight, transpose, precision)


def _gpacked(conv, transpose, precision):
    """Block-diagonal tensor-core tiles of a grouped 3x3 weight, re-made (in place: the address is baked into captured
    CUDA graphs) when the weight changed — same validation as ops.PackCache."""
    cache = conv.__dict__.setdefault("_b200lp_gpack", {})
    key = (transpose, precision)
    w = conv.weight
    stamp = ops._stamp(w)
    hit = cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    reuse = hit[1] if hit is not None and hit[1].device == w.device else None
    packed = K.pack_gconv_weight(w.detach(), transpose=transpose, precision=precision, out=reuse)
    cache[key] = (stamp, packed)
    return packed


def _stem_packed(conv):
    """The (64, 3, 7, 7) stem weight as a zero-padded (64, STEM_KP, 1, 1) 1x1 weight, packed for the bf16x3 GEMM."""
    st = conv.__dict__.get("_b200lp_stem")
    if st is None or st[0].device != conv.weight.device:
        pad = torch.zeros((64, K.STEM_KP, 1, 1), dtype=torch.float32, device=conv.weight.device)
        st = conv.__dict__["_b200lp_stem"] = [pad, None]
    st[0].view(64, K.STEM_KP)[:, :147].copy_(conv.weight.detach().reshape(64, 147))
    st[1] = K.pack_conv_weight(st[0], precision=K.BF16X3, out=st[1])
    return st[1]


def _bn_1x1(a_split, conv, bn, count):
    """raw = conv1x1(a) on the tensor cores (bf16x3), then this layer's BatchNorm statistics."""
    raw = K.conv_fwd(a_split, _packed(conv, False, K.BF16X3), 1)
    part = K.col_stats(raw.view(-1, raw.shape[-1])) if _stats_needed(bn) else None
    return raw, _BN(bn, part, count)


def blocks_of(net):
    return [blk for layer in (net.layer1, net.layer2, net.layer3, net.layer4) for blk in layer]


def forward(net, x_nchw, need_bwd):
    """x_nchw (N, 3, H, W) float32 CUDA -> (embeddings (N, num_classes), saved state for `backward` or None).
    Honours the modules' train / eval flags (batch statistics + running-statistics updates vs running statistics)."""
    x = x_nchw.contiguous().float()
    n, _, hh, ww = x.shape
    h, w = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    saved = {"blocks": []} if need_bwd else None

    # stem: 7x7/2 conv as a GEMM over the patch matrix, BatchNorm + ReLU folded into the max-pool's loads
    if need_bwd:
        col, cols = K.im2col7x7_s2(x, want_f32=True, want_split=True)
    else:
        col, cols = None, K.im2col7x7_s2(x, want_f32=False, want_split=True)
    r0 = K.conv_fwd(cols, _stem_packed(net.conv1), 1)
    del cols
    bn0 = _BN(net.bn1, K.col_stats(r0.view(-1, 64)) if _stats_needed(net.bn1) else None, n * h * w)
    # fp32 copy unrounded like every block output below (the TF32 weight-gradient GEMM reads it as is)
    a_f32, a_split, idx = K.maxpool3x3s2_fwd(r0, bn0.scale, bn0.shift, want_f32=True, want_split=True,
                                             want_idx=need_bwd, round_tf32=False)
    if need_bwd:
        saved.update(col=col, r0=r0, bn0=bn0, idx=idx, stem_hw=(h, w))
    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1

    blocks = blocks_of(net)
    for bi, blk in enumerate(blocks):
        last = bi == len(blocks) - 1
        s = blk.conv2.stride[0]
        ho, wo = h // s, w // s
        r1, bn1 = _bn_1x1(a_split, blk.conv1, blk.bn1, n * h * w)
        cpg = blk.conv2.weight.shape[1]
        a1_f32 = None
        if s == 1 and blk.conv2.weight.shape[0] % 64 == 0 and K.gconv_tensor_cores(n, h, w, blk.conv2.weight.shape[0], cpg):
            # stride-1 grouped conv on the tensor cores (bf16x3, block-diagonal 64-channel tiles): its operand relu(bn1(r1))
            # is materialised once — (hi, lo) planes for this GEMM and the tf32 copy the weight gradient reads later
            a1 = K.bn_act(r1, bn1.scale, bn1.shift, act=1, round_tf32=True, want_f32=need_bwd, want_split=True)
            a1_f32, a1_split = a1 if need_bwd else (None, a1)
            r2 = K.conv_fwd(a1_split, _gpacked(blk.conv2, False, K.BF16X3), 3, grouped=cpg)
            del a1_split
            part2 = K.col_stats(r2.view(-1, r2.shape[-1])) if _stats_needed(blk.bn2) else None
        elif _stats_needed(blk.bn2):
            r2, part2 = K.gconv3x3_fwd(r1, blk.conv2.weight.detach(), bn1.scale, bn1.shift, stride=s, want_stats=True)
        else:
            r2, part2 = K.gconv3x3_fwd(r1, blk.conv2.weight.detach(), bn1.scale, bn1.shift, stride=s), None
        bn2 = _BN(blk.bn2, part2, n * ho * wo)
        if need_bwd:
            a2_f32, a2_split = K.bn_act(r2, bn2.scale, bn2.shift, act=1, round_tf32=True, want_f32=True, want_split=True)
        else:
            a2_f32, a2_split = None, K.bn_act(r2, bn2.scale, bn2.shift, act=1, want_f32=False, want_split=True)
        r3, bn3 = _bn_1x1(a2_split, blk.conv3, blk.bn3, n * ho * wo)
        del a2_split
        rd = bnd = sub_f32 = None
        if blk.downsample is not None:
            if s == 2:
                sub_f32, sub_split = K.subsample2(a_f32 if need_bwd else None, a_split)
            else:
                sub_f32, sub_split = a_f32, a_split
            rd, bnd = _bn_1x1(sub_split, blk.downsample[0], blk.downsample[1], n * ho * wo)
            del sub_split
            res, res_sc, res_sh = rd, bnd.scale, bnd.shift
        else:
            res, res_sc, res_sh = a_f32, None, None
        # block output relu(bn3(r3) + identity): fp32 copy UNROUNDED (it is the next block's exact identity branch; the
        # TF32 weight-gradient GEMM reads it as is) + the (hi, lo) planes for the next block's forward GEMMs
        # + (for the backward pass) the block's ReLU mask as 4 bits per float4: both BatchNorm-backward passes of bn3 read
        # that byte instead of the 16-byte output
        outs = K.bn_act(r3, bn3.scale, bn3.shift, res=res, res_scale=res_sc, res_shift=res_sh, act=1, round_tf32=False,
                        want_f32=True, want_split=not last, want_mask=need_bwd)
        outs = outs if isinstance(outs, tuple) else (outs,)
        out_f32 = outs[0]
        out_split = outs[1] if not last else None
        out_mask = outs[-1] if need_bwd else None
        if need_bwd:
            saved["blocks"].append(dict(blk=blk, a_f32=a_f32, sub_f32=sub_f32, r1=r1, bn1=bn1, a1_f32=a1_f32, r2=r2, bn2=bn2, a2_f32=a2_f32,
                                        r3=r3, bn3=bn3, rd=rd, bnd=bnd, out=out_f32, out_mask=out_mask, hw=(h, w), stride=s))
        a_f32, a_split = out_f32, out_split
        h, w = ho, wo

    pooled = K.avgpool_fwd(a_f32)
    fc = net.fc
    emb = K.pw_conv(pooled, fc.weight.detach(), bias=fc.bias.detach() if fc.bias is not None else None)
    if need_bwd:
        saved.update(pooled=pooled, last_hw=(h, w))
    return emb, saved


class _Grads:
    """Parameter gradients of one backward pass: accumulated in place into the active gradient sinks
    (b200lp.ops.direct_grads — the runner's flat bucket) or collected as tensors for autograd."""

    def __init__(self, params):
        self.index = {id(p): i for i, p in enumerate(params)}
        self.out = [None] * len(params)

    def sink(self, p):
        return ops._sink(p)

    def give(self, p, g):
        """Hand a freshly computed gradient tensor (same shape as p) to autograd, or add it into the sink."""
        s = ops._sink(p)
        if s is not None:
            s.add_(g.view_as(s))
        else:
            i = self.index[id(p)]
            self.out[i] = g.view_as(p) if self.out[i] is None else self.out[i] + g.view_as(p)


def _wgrad_1x1(grads, conv, x_f32, dy):
    """Weight gradient of a 1x1 convolution on the TF32 tensor cores, into the sink or a fresh tensor."""
    if not conv.weight.requires_grad:
        return
    s = grads.sink(conv.weight)
    if s is not None:
        K.conv_wgrad_sn_acc(x_f32, dy, 1, s)
    else:
        g = torch.empty_like(conv.weight)
        K.conv_wgrad_sn_acc(x_f32, dy, 1, g, accumulate=False)
        grads.give(conv.weight, g)


def _bn_backward(grads, st, dy, x_raw, mask_mode, mask_src=None, round_tf32=False, want_dz=False):
    """BatchNorm (+ReLU) backward of the layer described by `st`; gamma / beta gradients go to the sinks or autograd."""
    bn = st.mod
    need = bn.weight.requires_grad
    sg, sb = (grads.sink(bn.weight), grads.sink(bn.bias)) if need else (None, None)
    if sg is not None and sb is not None:
        dx, _, _, dz = K.bn_bwd(dy, x_raw, st.mean, st.rstd, bn.weight.detach(), st.scale, st.shift, mask_src=mask_src,
                                mask_mode=mask_mode, dgamma=sg, dbeta=sb, accumulate=True, batch_stats=st.batch_stats,
                                round_tf32=round_tf32, want_dz=want_dz)
    else:
        dx, dg, db, dz = K.bn_bwd(dy, x_raw, st.mean, st.rstd, bn.weight.detach(), st.scale, st.shift, mask_src=mask_src,
                                  mask_mode=mask_mode, batch_stats=st.batch_stats, round_tf32=round_tf32, want_dz=want_dz)
        if need:
            grads.give(bn.weight, dg)
            grads.give(bn.bias, db)
    return dx, dz


TRACE = None     # tests: a list that receives (block record, incoming gradient, outgoing gradient) per block


def backward(net, saved, d_emb, params):
    """d_emb (N, num_classes) -> list of parameter gradients aligned with `params` (None where a sink took it)."""
    grads = _Grads(params)
    fc = net.fc
    d_emb = d_emb.contiguous().float()
    pooled = saved["pooled"]
    # classifier: dW = d_emb^T pooled, db = column sums, d_pooled = d_emb W
    if fc.weight.requires_grad:
        s = grads.sink(fc.weight)
        if s is not None:
            K.sgemm(d_emb, pooled, trans_a=True, acc_into=s)
        else:
            grads.give(fc.weight, K.sgemm(d_emb, pooled, trans_a=True))
    if fc.bias is not None and fc.bias.requires_grad:
        s = grads.sink(fc.bias)
        if s is not None:
            K.bias_grad(d_emb, acc_into=s)
        else:
            grads.give(fc.bias, K.bias_grad(d_emb))
    d_out = K.avgpool_bwd(K.sgemm(d_emb, fc.weight.detach()), saved["last_hw"])

    for rec in reversed(saved["blocks"]):
        blk, s = rec["blk"], rec["stride"]
        h, w = rec["hw"]
        # final ReLU + bn3 (dz3 = masked gradient, also the identity branch's gradient)
        d_block_out = d_out if TRACE is not None else None
        dr3, dz3 = _bn_backward(grads, rec["bn3"], d_out, rec["r3"], 4, mask_src=rec["out_mask"], round_tf32=True, want_dz=True)
        del d_out
        d_a2 = K.conv_fwd(dr3, _packed(blk.conv3, True, K.TF32), 1)
        _wgrad_1x1(grads, blk.conv3, rec["a2_f32"], dr3)
        del dr3
        w2 = blk.conv2.weight
        cpg = w2.shape[1]
        tc = K.gconv_tensor_cores(d_a2.shape[0], h, w, w2.shape[0], cpg)
        dr2, _ = _bn_backward(grads, rec["bn2"], d_a2, rec["r2"], 2, round_tf32=tc)
        del d_a2
        if tc:
            # tensor cores: stride 2 = the stride-1 kernels on the gradient placed at the even positions of the input grid
            if s == 2:
                dr2 = K.zero_stuff2(dr2)
            d_a1 = K.gconv3x3_dgrad(dr2, w2.detach(), (h, w), packed=_gpacked(blk.conv2, True, K.TF32))
        else:
            d_a1 = K.gconv3x3_dgrad(dr2, w2.detach(), (h, w), stride=s)
        if w2.requires_grad:
            sk = grads.sink(w2)
            if tc:
                a1 = rec["a1_f32"]        # kept by a tensor-core forward; else materialised here
                if a1 is None:
                    a1 = K.bn_act(rec["r1"], rec["bn1"].scale, rec["bn1"].shift, act=1, round_tf32=True, want_f32=True,
                                  want_split=False)
                g2 = K.gconv3x3_wgrad_tc(a1, dr2, cpg, acc_into=sk)
                del a1
            else:
                g2 = K.gconv3x3_wgrad(rec["r1"], dr2, cpg, rec["bn1"].scale, rec["bn1"].shift, stride=s, acc_into=sk)
            if sk is None:
                grads.give(w2, g2)
        del dr2
        dr1, _ = _bn_backward(grads, rec["bn1"], d_a1, rec["r1"], 2, round_tf32=True)
        del d_a1
        _wgrad_1x1(grads, blk.conv1, rec["a_f32"], dr1)
        wt1 = _packed(blk.conv1, True, K.TF32)
        if blk.downsample is not None:
            drd, _ = _bn_backward(grads, rec["bnd"], dz3, rec["rd"], 0, round_tf32=True)
            _wgrad_1x1(grads, blk.downsample[0], rec["sub_f32"], drd)
            d_sub = K.conv_fwd(drd, _packed(blk.downsample[0], True, K.TF32), 1)
            if s == 2:
                d_out = K.conv_fwd(dr1, wt1, 1)
                K.scatter_add2(d_sub, d_out)
            else:
                d_out = K.conv_fwd(dr1, wt1, 1, residual=d_sub, residual_mode=1)
        else:
            d_out = K.conv_fwd(dr1, wt1, 1, residual=dz3, residual_mode=1)
        if TRACE is not None:
            TRACE.append((rec, d_block_out, d_out))

    # stem: max-pool gather, bn1 + ReLU backward, weight gradient through the patch matrix
    d_act0 = K.maxpool3x3s2_bwd(d_out, saved["idx"], saved["stem_hw"])
    dr0, _ = _bn_backward(grads, saved["bn0"], d_act0, saved["r0"], 2, round_tf32=True)
    if net.conv1.weight.requires_grad:
        g = torch.empty((64, K.STEM_KP, 1, 1), dtype=torch.float32, device=dr0.device)
        K.conv_wgrad_sn_acc(saved["col"], dr0, 1, g, accumulate=False)
        grads.give(net.conv1.weight, g.view(64, K.STEM_KP)[:, :147].reshape(64, 3, 7, 7))
    return grads.out


class ResNeXtFn(torch.autograd.Function):
    """embeddings = resnext(x) as ONE autograd node over the module's parameters."""

    @staticmethod
    def forward(ctx, net, x, *params):
        need_bwd = any(ctx.needs_input_grad[2:])
        emb, saved = forward(net, x, need_bwd)
        ctx.net, ctx.saved, ctx.params = net, saved, params
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        if ctx.saved is None:
            raise RuntimeError("ResNeXtFn.backward without saved state")
        out = backward(ctx.net, ctx.saved, d_emb, ctx.params)
        return (None, None) + tuple(out)


def apply(net, x_nchw):
    """Differentiable (w.r.t. the parameters) forward of `net` on `x_nchw` through the kernel schedule."""
    params = tuple(net.parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return ResNeXtFn.apply(net, x_nchw, *params)
    emb, _ = forward(net, x_nchw, False)
    return emb
