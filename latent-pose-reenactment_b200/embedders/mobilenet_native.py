"""Forward pass of a torchvision MobileNetV2 module as a schedule of libb200lp kernels (csrc/mobilenet.cu).

The module keeps owning every parameter and buffer (checkpoint keys `pose_encoder.*` are torchvision's); this file
only sequences kernels over them: conv -> (batch statistics) -> bn_finalize, with every BatchNorm + ReLU6 applied by
the CONSUMER kernel on load.  Same results as `net(x)` (reference: embedders/unsupervised_pose_separate_embResNeXt_
segmentation.py:56-58) including the train-mode running-statistics updates; used whenever no gradient is needed
through the encoder (drive.py, fine-tuning steps, EMA forward) — the differentiable path (meta-training) stays on
the torch modules.
"""
import torch

from b200lp import kernels as K


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


def forward(net, x_nchw):
    """x_nchw: (N, 3, H, W) float32 CUDA -> (N, num_classes).  Honours net.training (batch statistics + running-stat
    updates + dropout) vs eval (running statistics)."""
    x = x_nchw.contiguous().float()
    n = x.shape[0]
    f = net.features

    def finalize(bn, part, count):
        training = bn.training or not bn.track_running_stats
        return K.bn_finalize(bn, part if training else None, count, training)

    def stats_on(bn):
        return bn.training or not bn.track_running_stats

    # stem: conv3x3 s2 -> lazy (raw, scale, shift): consumers apply BN + ReLU6 on load
    conv, bn = _conv_bn(f[0])
    if stats_on(bn):
        raw, part = K.mbv2_stem(x, conv.weight.detach(), want_stats=True)
    else:
        raw, part = K.mbv2_stem(x, conv.weight.detach()), None
    sc, sh = finalize(bn, part, raw.numel() // raw.shape[-1])
    cur, cur_sc, cur_sh = raw, sc, sh           # lazy activation relu6(cur*sc+sh); materialised when cur_sc is None

    def pw(inp, inp_sc, inp_sh, relu6, conv, bn):
        nn_, h, w, c = inp.shape
        wt = conv.weight.detach().reshape(conv.out_channels, conv.in_channels)
        if stats_on(bn):
            y, part = K.pw_conv(inp.reshape(-1, c), wt, inp_sc, inp_sh, relu6, want_stats=True)
        else:
            y, part = K.pw_conv(inp.reshape(-1, c), wt, inp_sc, inp_sh, relu6), None
        sc_, sh_ = finalize(bn, part, y.shape[0])
        return y.reshape(nn_, h, w, conv.out_channels), sc_, sh_

    for blk in list(f)[1:-1]:
        layers = list(blk.conv)
        block_in = cur if cur_sc is None else None      # materialised block input (needed for the skip connection)
        if len(layers) == 4:                            # expand 1x1 + BN + ReLU6
            conv, bn = _conv_bn(layers[0])
            e, e_sc, e_sh = pw(cur, cur_sc, cur_sh, cur_sc is not None, conv, bn)
            dw_seq = layers[1]
        else:                                           # t = 1 block: depthwise acts on the incoming lazy activation
            assert cur_sc is not None, "depthwise conv needs a BatchNorm+ReLU6 producer"
            e, e_sc, e_sh = cur, cur_sc, cur_sh
            dw_seq = layers[0]
        conv, bn = _conv_bn(dw_seq)
        assert conv.groups == conv.in_channels == conv.out_channels and conv.kernel_size == (3, 3)
        if stats_on(bn):
            d, part = K.dw_conv3x3(e, conv.weight.detach(), e_sc, e_sh, conv.stride[0], want_stats=True)
        else:
            d, part = K.dw_conv3x3(e, conv.weight.detach(), e_sc, e_sh, conv.stride[0]), None
        d_sc, d_sh = finalize(bn, part, d.numel() // d.shape[-1])
        conv, bn = layers[-2], layers[-1]               # linear 1x1 + BN
        p, p_sc, p_sh = pw(d, d_sc, d_sh, True, conv, bn)
        res = None
        if blk.use_res_connect:
            assert block_in is not None
            res = block_in
        cur, cur_sc, cur_sh = K.bn_apply(p, p_sc, p_sh, residual=res), None, None

    conv, bn = _conv_bn(f[-1])                          # 1x1 -> 1280, BN, ReLU6, global average pool
    last, l_sc, l_sh = pw(cur, cur_sc, cur_sh, cur_sc is not None, conv, bn)
    pooled = K.bn_relu6_avgpool(last, l_sc, l_sh)
    drop, lin = net.classifier[0], net.classifier[1]
    if net.training and isinstance(drop, torch.nn.Dropout) and drop.p > 0:
        pooled = torch.nn.functional.dropout(pooled, drop.p, True)     # RNG-dependent, (N, 1280): stays a torch op
    return K.pw_conv(pooled, lin.weight.detach(), bias=lin.bias.detach() if lin.bias is not None else None)
