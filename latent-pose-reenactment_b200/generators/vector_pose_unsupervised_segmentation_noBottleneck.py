"""B200-native generator plugin — drop-in for the reference's
`generators/vector_pose_unsupervised_segmentation_noBottleneck.py` (Wrapper :8-28, Generator :40-181).

Same plugin API (`Wrapper.get_args/get_net`, `forward(data_dict)`, `enable_finetuning`), same state_dict keys and
shapes (checkpoints interchange), same parameter order (optimizer state interchanges).  The arithmetic is a
schedule of libb200lp kernels: per decoder block  adain_relu -> tcgen05 conv -> adain_relu -> tcgen05 conv(+skip),
with NHWC activations internally and NCHW images at the `data_dict` boundary.
"""
import math

import torch
from torch import nn

from b200lp import lib as b200lp_lib
from b200lp import ops
from generators.common import blocks


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--gen_constant_input_size', type=int, default=4)
        parser.add('--gen_num_residual_blocks', type=int, default=2)
        parser.add('--gen_padding', type=str, default='zero', help='zero|reflection')
        parser.add('--norm_layer', type=str, default='in')

    @staticmethod
    def get_net(args):
        if 'gen_constant_input_size' not in args:   # old checkpoints (reference :19-21)
            args.gen_constant_input_size = 4
        net = Generator(
            args.gen_padding, args.in_channels, args.out_channels + 1,
            args.num_channels, args.max_num_channels, args.embed_channels, args.pose_embedding_size,
            args.norm_layer, args.gen_constant_input_size, args.gen_num_residual_blocks,
            args.image_size)
        return net.to(args.device)


class Constant(nn.Module):
    """Learned (1, C, s, s) input of the decoder (reference :31-37); stored NCHW for checkpoint compatibility."""

    def __init__(self, *shape):
        super().__init__()
        self.constant = nn.Parameter(torch.ones(1, *shape))


class Generator(nn.Module):
    def __init__(self, padding, in_channels, out_channels, num_channels, max_num_channels, identity_embedding_size,
                 pose_embedding_size, norm_layer, gen_constant_input_size, gen_num_residual_blocks,
                 output_image_size):
        super().__init__()
        if padding != 'zero':
            if padding == 'reflection':
                raise NotImplementedError("B200 generator: only `gen_padding=zero` (the shipped configs) is native")
            raise Exception('Incorrect `padding` argument, required `zero` or `reflection`')
        if norm_layer != 'in':
            raise NotImplementedError("B200 generator: only `norm_layer=in` (AdaIN over InstanceNorm) is native")
        if out_channels != 4:
            raise NotImplementedError("B200 generator: the fused tail kernel emits RGB + segmentation (4 channels)")
        assert math.log2(output_image_size / gen_constant_input_size).is_integer(), \
            "`gen_constant_input_size` must be `image_size` divided by a power of 2"
        num_up = int(math.log2(output_image_size / gen_constant_input_size))
        nonclamped = num_channels * (2 ** num_up)
        c = min(nonclamped, max_num_channels)

        self.constant = Constant(c, gen_constant_input_size, gen_constant_input_size)

        children = {}
        self.adain_sizes = []
        idx = 0
        for _ in range(gen_num_residual_blocks):
            children[str(idx)] = blocks.AdaResBlock(c, c, upsample=False)
            self.adain_sizes += [c, c]
            idx += 1
        for _ in range(num_up):
            cin = c
            nonclamped //= 2
            c = min(nonclamped, max_num_channels)
            children[str(idx)] = blocks.AdaResBlock(cin, c, upsample=True)
            self.adain_sizes += [cin, c]
            idx += 1
        self.num_blocks = idx
        self.adain_sizes.append(c)           # final AdaIN (reference decoder_blocks[num_blocks])
        # reference Sequential: [blocks..., AdaIN, ReLU, conv, Tanh] -> the conv sits at index num_blocks + 2
        children[str(idx + 2)] = blocks.SNConv(c, out_channels, 3, bias=True)
        self.decoder_blocks = blocks.Slots(**children)

        self.identity_embedding_size = identity_embedding_size
        self.pose_embedding_size = pose_embedding_size
        joint = identity_embedding_size + pose_embedding_size
        hidden = max(joint, 512)
        self.affine_params_projector = blocks.Slots(**{
            # torch's default spectral_norm eps here (reference :97-101 passes none); 1e-4 is only used for the convs
            "0": blocks.SNLinear(joint, hidden, eps=1e-12),
            "2": blocks.SNLinear(hidden, self.get_num_affine_params(), eps=1e-12)})
        self.finetuning = False
        # tensor-core operand precision of the decoder convolutions: 'bf16x3' (default; three bf16 MMAs per K step on
        # (hi, lo) operand planes, ~fp32 accuracy, generator RGB within 1e-3 of the fp32 reference) or 'tf32' (one MMA
        # per K step, 1.5x less tensor time, ~3e-3 max-abs on nets with O(1) AdaIN gains).  Not part of the state_dict.
        self.precision = 'bf16x3'
        # final AdaIN + conv3x3(C -> 4) + tanh tail: tcgen05 bf16x3 conv on a weight padded to 32 outputs (True) or the
        # fp32 CUDA-core kernel (False; also taken for shapes the tensor-core kernel does not cover)
        self.tail_on_tensor_cores = True

    def get_num_affine_params(self):
        return sum(2 * c for c in self.adain_sizes)

    def compute_affine_params(self, data_dict):
        """assign_embeddings (reference :127-137): [identity | pose] -> SN-Linear -> ReLU -> SN-Linear."""
        if self.finetuning:
            identity = self.identity_embedding.expand(len(data_dict['pose_embedding']), -1)
        else:
            identity = data_dict['embeds']
        joint = torch.cat((identity, data_dict['pose_embedding']), dim=1)
        h = torch.relu(self.affine_params_projector.slot(0)(joint))
        return self.affine_params_projector.slot(2)(h)

    def enable_finetuning(self, data_dict=None):
        """Reference :139-163: the identity embedding becomes a trainable parameter `identity_embedding` (1, E)."""
        if data_dict is None:
            some_parameter = next(iter(self.parameters()))
            identity_embedding = torch.rand(1, self.identity_embedding_size).to(some_parameter)
        else:
            identity_embedding = data_dict['embeds']
        if self.finetuning:
            with torch.no_grad():
                self.identity_embedding.copy_(identity_embedding)
        else:
            self.identity_embedding = nn.Parameter(identity_embedding)
            self.finetuning = True

    def forward(self, data_dict):
        b200lp_lib.require_device()
        # one batched power iteration / sigma evaluation for all 25 spectral-normalised weights of the generator
        # (22 decoder convs, the tail conv, the two projector layers): 3 launches instead of ~14 per weight
        convs = []
        for i in range(self.num_blocks):
            convs += self.decoder_blocks.slot(i).tensor_core_convs()
        tail = self.decoder_blocks.slot(self.num_blocks + 2)
        blocks.spectral_sigmas(convs + [tail, self.affine_params_projector.slot(0), self.affine_params_projector.slot(2)])
        affine = self.compute_affine_params(data_dict)        # (B, sum 2C): per AdaIN [beta | gamma]
        batch = affine.shape[0]
        # all 17 (gamma, beta) pairs from one autograd node (one concatenation in backward instead of 34 slice backwards)
        pairs = iter(ops.split_affine(affine, self.adain_sizes))

        def take(c):
            gamma, beta = next(pairs)
            assert gamma.shape[1] == c
            return gamma, beta

        # constant (1,C,s,s) NCHW parameter -> (B,s,s,C) NHWC
        x = self.constant.constant.permute(0, 2, 3, 1).expand(batch, -1, -1, -1).contiguous()
        x_split = None
        for i in range(self.num_blocks):
            blk = self.decoder_blocks.slot(i)
            g0, b0 = take(blk.in_channels)
            g1, b1 = take(blk.out_channels)
            # the next block's 1x1 skip conv reads this output as a tensor-core operand: produce it in operand form here
            nxt = self.decoder_blocks.slot(i + 1) if i + 1 < self.num_blocks else None
            feeds_skip_conv = nxt is not None and nxt.skip is not None
            x, x_split = blk(x, g0, b0, g1, b1, feeds_skip_conv, x_split=x_split, precision=self.precision)
        g, bt = take(self.adain_sizes[-1])
        if getattr(self, 'tail_on_tensor_cores', True) and ops.tail_tensor_core_ok(x, tail.weight_orig):
            fake_rgbs, fake_segm = ops.adain_tail(x, g, bt, tail.weight_orig, tail.scale(), tail.bias)
        else:
            a = ops.adain_relu(x, g, bt, round_out=False)
            fake_rgbs, fake_segm = ops.gen_tail(a, tail.weight_orig, tail.scale(), tail.bias)
        data_dict['fake_rgbs'] = fake_rgbs
        data_dict['fake_segm'] = fake_segm
