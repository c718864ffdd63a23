"""Shared building blocks of the B200-native generator / discriminator plugins.

Host-side mirror of the reference's `generators/common/blocks.py` (AdaptiveNorm2d :6-26, ResBlock :47-111) for the
two configurations the shipped configs instantiate: `norm_layer='adain'` (generator) and `norm_layer='none'`
(discriminator).  Modules here only *own parameters* (with the reference's state_dict names: `weight_orig`,
`weight_u`, `weight_v`, `bias`) and sequence kernel calls from `b200lp.ops`; activations are NHWC.
"""
import math

import torch
from torch import nn

from b200lp import ops


class Slots(nn.Module):
    """A container whose children are registered under explicit (numeric) names, so that state_dict keys line up
    with the reference's nn.Sequential indices even though the parameter-free layers in between do not exist here."""

    def __init__(self, **children):
        super().__init__()
        for name, child in children.items():
            self.add_module(name, child)

    def slot(self, index):
        return self._modules[str(index)]


class SpectralNormed(nn.Module):
    """Owner of one spectral-normalised weight: parameters `bias` (optional, registered first like the reference's
    nn.Conv2d/nn.Linear after torch.nn.utils.spectral_norm re-registers `weight_orig`), `weight_orig`; buffers
    `weight_u`, `weight_v`.

    `inv_sigma()` follows torch's SpectralNorm.compute_weight: in training mode one in-place power iteration on the
    buffers, then sigma = u^T W v with u, v treated as constants; returns 1/sigma as a 1-element tensor attached to
    `weight_orig`'s autograd graph.  The kernels consume (weight_orig, 1/sigma) directly — W/sigma is never
    materialised.
    """

    def __init__(self, weight_shape, bias, eps=1e-4):
        super().__init__()
        out_features = weight_shape[0]
        fan_in = int(math.prod(weight_shape[1:]))
        if bias:
            bound = 1 / math.sqrt(fan_in)
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.bias = None
        w = torch.empty(*weight_shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight_orig = nn.Parameter(w)
        self.eps = eps
        u = torch.randn(out_features)
        v = torch.randn(fan_in)
        self.register_buffer("weight_u", u / u.norm().clamp_min(eps))
        self.register_buffer("weight_v", v / v.norm().clamp_min(eps))
        self.pack_cache = ops.PackCache()   # tensor-core layouts of weight_orig, re-made when the weight changes
        self._sn_scratch = None             # device scratch of the batched power iteration (lazily allocated)
        self._pre = None                    # (1/sigma, (u, v) snapshots) deposited by spectral_sigmas() for the next use

    def inv_sigma(self):
        w = self.weight_orig
        wm = w.reshape(w.shape[0], -1)
        if self.training:
            with torch.no_grad():
                v = torch.mv(wm.t(), self.weight_u)
                v = v / v.norm().clamp_min(self.eps)
                u = torch.mv(wm, v)
                u = u / u.norm().clamp_min(self.eps)
                self.weight_v.copy_(v)
                self.weight_u.copy_(u)
        u = self.weight_u.clone()
        v = self.weight_v.clone()
        sigma = torch.dot(u, torch.mv(wm, v))
        return (1.0 / sigma).reshape(1)

    def scale(self, detach=False):
        """1/sigma WITH its autograd edge to `weight_orig`, for consumers that differentiate through plain autograd
        (linear / embedding layers, Cin=3 stems, generator tail): the batched kernel result deposited by
        spectral_sigmas() wrapped in ops.InvSigmaFn, else torch's per-layer formula `inv_sigma()`."""
        if self._pre is not None:
            s, (su, sv) = self._pre
            self._pre = None
            if detach or not (torch.is_grad_enabled() and self.weight_orig.requires_grad):
                return s
            return ops.inv_sigma_edge(self.weight_orig, s, su, sv)
        s = self.inv_sigma()
        return s.detach() if detach else s

    def operands_edge(self, detach=False):
        """(weight_orig, 1/sigma with autograd edge, bias) — see scale()."""
        s = self.scale(detach)
        if detach:
            return self.weight_orig.detach(), s, (self.bias.detach() if self.bias is not None else None)
        return self.weight_orig, s, self.bias

    def operands(self, detach=False):
        """(weight_orig, 1/sigma, bias, extras) as the kernels consume them; extras = dict(cache=PackCache, sn=(u, v) or
        None).  If spectral_sigmas() deposited a batched result for this layer it is consumed here (sigma then has no
        autograd edge; the rank-1 gradient term is applied by the kernel using the (u, v) snapshots); otherwise the
        per-layer autograd path `inv_sigma()` is used.  `detach=True` cuts the parameter gradients (the discriminator
        pass whose weight gradients the training step discards anyway); the power iteration still runs."""
        if self._pre is not None:
            s, sn = self._pre
            self._pre = None
        else:
            s, sn = self.inv_sigma(), None
        extras = dict(cache=self.pack_cache, sn=sn)
        if detach:
            return (self.weight_orig.detach(), s.detach(), (self.bias.detach() if self.bias is not None else None),
                    extras)
        return self.weight_orig, s, self.bias, extras


def spectral_sigmas(layers):
    """One batched spectral-norm pass (3 kernel launches) for a list of SpectralNormed conv layers: power iteration in
    training mode (buffers updated in place), 1/sigma and (u, v) snapshots deposited on each layer for its next
    `operands()` call.  Replaces len(layers) x ~14 tiny launches of torch's spectral_norm hooks per network pass."""
    from b200lp import kernels as K
    from b200lp import lib as L
    if not layers:
        return
    max_t = L.load().b200lp_sn_max_tensors()
    training = layers[0].training
    with torch.no_grad():
        for i in range(0, len(layers), max_t):
            chunk = layers[i:i + max_t]
            packed = []
            for m in chunk:
                w = m.weight_orig
                if m._sn_scratch is None or m._sn_scratch.device != w.device:
                    m._sn_scratch = K.sn_scratch(w)
                packed.append((w.detach(), m.weight_u, m.weight_v, float(m.eps), m._sn_scratch))
            inv, snaps = K.sn_sigma_multi(packed, training)
            for j, m in enumerate(chunk):
                m._pre = (inv[j:j + 1], snaps[j])


class SNConv(SpectralNormed):
    def __init__(self, in_channels, out_channels, ksize, bias, eps=1e-4):
        super().__init__((out_channels, in_channels, ksize, ksize), bias, eps)
        self.ksize = ksize


class SNLinear(SpectralNormed):
    def __init__(self, in_features, out_features, eps=1e-4):
        super().__init__((out_features, in_features), True, eps)

    def forward(self, x):
        # (x W^T) / sigma + b  ==  F.linear(x, W / sigma, b) without materialising W / sigma
        if x.dim() == 2:
            return ops.linear(x, self.weight_orig, self.scale(), self.bias)       # one SGEMM kernel, scale + bias fused
        return torch.nn.functional.linear(x, self.weight_orig) * self.scale() + self.bias


class AdaResBlock(nn.Module):
    """ResBlock(norm_layer='adain') as built by the generator's get_res_block / get_up_block
    (reference blocks.py:47-111, generator :45-51).  Children mirror the reference's Sequential indices:
    block.{3,7} (plain) or block.{4,8} (upsampling), skip.1.

        main: AdaIN -> ReLU -> [nearest 2x] -> conv3x3 -> AdaIN -> ReLU -> conv3x3      skip: [2x] -> conv1x1(+bias)

    Kernel schedule (NHWC): adain_relu(+2x, tf32)  ->  conv  ->  adain_relu  ->  conv (+ residual in the epilogue).
    The 1x1 skip conv is evaluated at the LOW resolution and nearest-upsampled inside the second conv's epilogue
    (conv1x1 and nearest-upsample commute: 4x fewer MACs, identical values).
    """

    def __init__(self, in_channels, out_channels, upsample):
        super().__init__()
        i0, i1 = (4, 8) if upsample else (3, 7)
        self.i0, self.i1 = i0, i1
        self.in_channels, self.out_channels, self.upsample = in_channels, out_channels, upsample
        self.block = Slots(**{str(i0): SNConv(in_channels, out_channels, 3, bias=False),
                              str(i1): SNConv(out_channels, out_channels, 3, bias=False)})
        self.skip = None
        if in_channels != out_channels or upsample:
            self.skip = Slots(**{"1": SNConv(in_channels, out_channels, 1, bias=True)})

    def forward(self, x, gamma0, beta0, gamma1, beta1, feeds_skip_conv, x_split=None, precision='bf16x3'):
        """x: block input (NHWC fp32).  `feeds_skip_conv`: the NEXT block applies a 1x1 conv to this block's output,
        i.e. the output is itself a tensor-core operand.  Returns (out, out_split) — out_split (bf16 hi/lo planes of
        `out`) only in bf16x3 precision when feeds_skip_conv, else None.

        precision 'tf32'  : adain_relu(tf32) -> conv -> adain_relu(tf32) -> conv(+skip), one TF32 MMA per K step;
        precision 'bf16x3': the same schedule with (hi, lo) bf16 operand planes and three MMAs per K step — generator
                            output within 1e-3 of the fp32 reference for O(1) AdaIN gains."""
        import os
        convs = [self.block.slot(self.i0), self.block.slot(self.i1)] + ([self.skip.slot(1)] if self.skip is not None else [])
        if precision == 'bf16x3' and not os.environ.get('B200LP_NO_GEN_BLOCK_NODE') and \
                all(m._pre is not None and m._pre[1] is not None for m in convs):
            # the whole block as one autograd node (ops.AdaResBlockFn): kernel-merged input gradients, tf32-rounded
            # gradient operands
            w0, s0, _, e0 = convs[0].operands()
            w1, s1, _, e1 = convs[1].operands()
            spec = dict(c0=(w0, s0, e0["cache"], e0["sn"]), c1=(w1, s1, e1["cache"], e1["sn"]), sk=None)
            if self.skip is not None:
                ws, ss, bs, es = convs[2].operands()
                spec["sk"] = (ws, ss, bs, es["cache"], es["sn"])
            out = ops.ada_res_block(x, gamma0, beta0, gamma1, beta1, x_split if self.skip is not None else None, spec,
                                    self.upsample, feeds_skip_conv)
            return out if feeds_skip_conv else (out, None)
        w0, s0, _, e0 = self.block.slot(self.i0).operands()
        w1, s1, _, e1 = self.block.slot(self.i1).operands()
        if precision == 'bf16x3':
            y1 = ops.adain_conv(x, gamma0, beta0, w0, s0, upsample2=self.upsample, **e0)
            if self.skip is not None:
                ws, ss, bs, es = self.skip.slot(1).operands()
                s = ops.conv2d(x, ws, ss, bias=bs, ksize=1, x_split=x_split, **es)
                mode = 2 if self.upsample else 1
            else:
                s, mode = x, 1
            out = ops.adain_conv(y1, gamma1, beta1, w1, s1, residual=s, residual_mode=mode,
                                 emit_split=feeds_skip_conv, **e1)
            return out if feeds_skip_conv else (out, None)
        a0 = ops.adain_relu(x, gamma0, beta0, upsample2=self.upsample)
        y1 = ops.conv2d(a0, w0, s0, ksize=3, **e0)
        a1 = ops.adain_relu(y1, gamma1, beta1)
        if self.skip is not None:
            ws, ss, bs, es = self.skip.slot(1).operands()
            s = ops.conv2d(x, ws, ss, bias=bs, ksize=1, **es)
            mode = 2 if self.upsample else 1
        else:
            s, mode = x, 1
        return ops.conv2d(a1, w1, s1, residual=s, residual_mode=mode, ksize=3, round_out=feeds_skip_conv, **e1), None

    def tensor_core_convs(self):
        return [m for m in (self.block.slot(self.i0), self.block.slot(self.i1),
                            self.skip.slot(1) if self.skip is not None else None) if m is not None]


class PlainResBlock(nn.Module):
    """ResBlock(norm_layer='none') as built by the discriminator (reference blocks.py:47-111 via
    discriminators/no_landmarks.py:37-43): children block.{2,5} (+bias), skip.0 (+bias).

        r = relu(x) [in place in the reference: the skip branch and the stored feature both see r]
        main: conv3x3(r)+b -> ReLU -> conv3x3+b -> [avgpool2]        skip: conv1x1(r)+b -> [avgpool2]

    Kernel schedule: conv(+bias, relu, tf32) -> conv(+bias) -> avgpool2(+skip) with the 1x1 skip conv evaluated on
    the pooled input (avg-pool and conv1x1 commute).
    """

    def __init__(self, in_channels, out_channels, downsample):
        super().__init__()
        self.in_channels, self.out_channels, self.downsample = in_channels, out_channels, downsample
        self.block = Slots(**{"2": SNConv(in_channels, out_channels, 3, bias=True),
                              "5": SNConv(out_channels, out_channels, 3, bias=True)})
        self.skip = None
        if in_channels != out_channels or downsample:
            self.skip = Slots(**{"0": SNConv(in_channels, out_channels, 1, bias=True)})

    def forward(self, r, detach_params=False):
        """r = tf32(relu(block input)).  Returns the block output (pre-ReLU)."""
        w0, s0, b0, e0 = self.block.slot(2).operands(detach_params)
        h = ops.conv2d(r, w0, s0, bias=b0, ksize=3, relu=True, round_out=True, **e0)
        w1, s1, b1, e1 = self.block.slot(5).operands(detach_params)
        if self.skip is not None:
            ws, ss, bs, es = self.skip.slot(0).operands(detach_params)
            rs = ops.avgpool2(r, None, round_out=True) if self.downsample else r
            s = ops.conv2d(rs, ws, ss, bias=bs, ksize=1, **es)
        else:
            s = r
        if self.downsample:
            h2 = ops.conv2d(h, w1, s1, bias=b1, ksize=3, **e1)
            return ops.avgpool2(h2, s)
        return ops.conv2d(h, w1, s1, bias=b1, residual=s, residual_mode=1, ksize=3, **e1)

    def tensor_core_convs(self):
        return [m for m in (self.block.slot(2), self.block.slot(5), self.skip.slot(0) if self.skip is not None else None)
                if m is not None]
