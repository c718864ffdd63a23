"""VGG feature ("perceptual") loss on libb200lp kernels — drop-in for the reference's
`criterions/common/perceptual_loss.py` (PerceptualLoss :19-110).

Same constructor contract (weight, vgg_weights_dir, net in {'caffe','face','pytorch'}), same weight files
(`vgg19-d01eb7cb.pth`: full torchvision-vgg19 state dict with the caffe classifier naming;
`vgg_face_weights.pth`: `vgg16().features` state dict), same arithmetic: `(x+1)/2`, mean/std normalisation,
features[0:30] with MaxPool->AvgPool, mean-L1 after each of the 13 ReLUs, target detached, weights frozen.

Execution: the frozen weights are packed once into the tensor-core layout; fake and real images run layer by layer
through direct-conv (Cin=3 stem, input normalisation fused) / tcgen05 implicit-GEMM convs (bias+ReLU+tf32 epilogue)
/ avg-pool kernels, with one L1 reduction kernel per tap; backward is the data-gradient chain only.
"""
import os

import torch
from torch import nn

from b200lp import kernels as K
from b200lp import ops

# torchvision layer indices inside `features` (first 30 layers)
_PLANS = {
    'vgg19': dict(convs=(0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28), pools=(4, 9, 18, 27)),
    'vgg16': dict(convs=(0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28), pools=(4, 9, 16, 23)),
}


class PerceptualLoss(nn.Module):
    def __init__(self, weight, vgg_weights_dir, net='caffe', normalize_grad=False):
        super().__init__()
        self.weight = weight
        if normalize_grad:
            raise NotImplementedError("normalize_grad is a no-op branch in the reference (:104-106) and is not used")
        num_layers = 30
        if net == 'caffe':
            arch = 'vgg19'
            sd = torch.load(os.path.join(vgg_weights_dir, 'vgg19-d01eb7cb.pth'), map_location='cpu')
            sd = {k[len('features.'):]: v for k, v in sd.items() if k.startswith('features.')}
            mean = torch.tensor([103.939, 116.779, 123.680]) / 255.
            std = torch.tensor([1., 1., 1.]) / 255.
        elif net == 'face':
            arch = 'vgg16'
            sd = torch.load(os.path.join(vgg_weights_dir, 'vgg_face_weights.pth'), map_location='cpu')
            mean = torch.tensor([103.939, 116.779, 123.680]) / 255.
            std = torch.tensor([1., 1., 1.]) / 255.
        elif net == 'pytorch':
            raise NotImplementedError("net='pytorch' needs the torchvision download (no network); shipped configs "
                                      "use 'caffe' and 'face'")
        else:
            raise ValueError(f"Unknown type of PerceptualLoss: expected '{{pytorch,caffe,face}}', got '{net}'")

        plan = _PLANS[arch]
        self.arch = arch
        self.plan = []
        for i in range(num_layers):
            if i in plan['convs']:
                self.plan.append(('conv0' if i == 0 else 'conv', i))
                self.register_buffer(f'w{i}', sd[f'{i}.weight'].float().contiguous(), persistent=False)
                self.register_buffer(f'b{i}', sd[f'{i}.bias'].float().contiguous(), persistent=False)
            elif i in plan['pools']:
                self.plan.append(('pool', i))
        self.register_buffer('mean', mean[None, :, None, None])
        self.register_buffer('std', std[None, :, None, None])
        self._packed = None

    def _apply(self, fn, *args, **kwargs):   # .to(device) / .cuda(): drop packed copies
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def pack(self):
        """Frozen weights -> tensor-core layouts (once per device placement)."""
        if self._packed is not None:
            return self._packed
        mean = self.mean.flatten()
        std = self.std.flatten()
        # ((x+1)/2 - mean)/std  ==  x * (0.5/std) + (0.5 - mean)/std        (perceptual_loss.py:88-98)
        packed = dict(plan=self.plan, wp={}, wpt={}, bias={},
                      pre_scale=(0.5 / std).contiguous(), pre_shift=((0.5 - mean) / std).contiguous())
        for kind, i in self.plan:
            w, b = getattr(self, f'w{i}', None), getattr(self, f'b{i}', None)
            if kind == 'conv0':
                packed['w0'], packed['b0'] = w, b
                packed['w0t'] = K.c3_transposed_weight(w)        # for the tensor-core data gradient of the stem
            elif kind == 'conv':
                packed['wp'][i] = K.pack_conv_weight(w, None, transpose=False)
                packed['wpt'][i] = K.pack_conv_weight(w, None, transpose=True)
                packed['bias'][i] = b
        self._packed = packed
        return packed

    def forward(self, input, target):
        return ops.vgg_perceptual(input, target, self.pack(), self.weight)
