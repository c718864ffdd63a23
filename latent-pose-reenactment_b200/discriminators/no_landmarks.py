"""B200-native projection discriminator — drop-in for the reference's `discriminators/no_landmarks.py`
(Wrapper :11-28, Discriminator :31-166): same plugin API, state_dict keys/shapes, parameter order, data_dict contract.

Three passes per step share weights (fake-for-G, fake.detach()-for-D, real); in training mode every pass runs one
more spectral-norm power iteration, like the reference's forward pre-hooks.  Feature maps are returned as
(B, C, H, W)-shaped tensors whose memory is NHWC (channels_last strides), so the feature-matching criterion reads
them without a layout copy.
"""
import math
import os

import torch
from torch import nn

from b200lp import lib as b200lp_lib
from b200lp import ops
from generators.common import blocks

from utils.radam import RAdam

torch.optim.RAdam = RAdam   # the reference swaps in its vendored RAdam at import time (:5-6)


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--dis_padding', type=str, default='zero', help='zero|reflection')
        parser.add('--dis_num_blocks', type=int, default=7)
        parser.add('--lr_dis', type=float, default=2e-4)

    @staticmethod
    def get_net(args):
        net = Discriminator(args.dis_padding, args.in_channels, args.out_channels, args.num_channels,
                            args.max_num_channels, args.embed_channels, args.dis_num_blocks, args.image_size,
                            args.num_labels).to(args.device)
        return net

    @staticmethod
    def get_optimizer(discriminator, args):
        from runners.holycow import optimizer_class      # Adam / RAdam, fused multi-tensor kernel on CUDA
        Optimizer = optimizer_class(args.optimizer, args.device)
        return Optimizer(discriminator.parameters(), lr=args.lr_dis, betas=(args.beta1, 0.999), eps=1e-5)


class SNEmbedding(blocks.SpectralNormed):
    """spectral_norm(nn.Embedding) (reference :84-86, fine-tuned variant :132-134 with torch's default eps 1e-12)."""

    def __init__(self, num_labels, channels, eps=1e-4):
        super().__init__((num_labels, channels), bias=False, eps=eps)
        with torch.no_grad():
            self.weight_orig.uniform_(-0.1, 0.1)

    def forward(self, label):
        if self.weight_orig.is_cuda:
            blocks.spectral_sigmas([self])      # batched kernels (3 launches) instead of torch's ~14 per call
        return torch.nn.functional.embedding(label, self.weight_orig) * self.scale()


class Discriminator(nn.Module):
    def __init__(self, padding, in_channels, out_channels, num_channels, max_num_channels, embed_channels,
                 dis_num_blocks, image_size, num_labels):
        super().__init__()
        if padding != 'zero':
            raise NotImplementedError("B200 discriminator: only `dis_padding=zero` (the shipped configs) is native")
        if in_channels != 3:
            raise NotImplementedError("B200 discriminator: the stem kernels take 3-channel images")
        self.out_channels = embed_channels

        self.down_block = blocks.Slots(**{"0": blocks.SNConv(in_channels, num_channels, 3, bias=True),
                                          "2": blocks.SNConv(num_channels, num_channels, 3, bias=True)})
        self.skip = blocks.Slots(**{"0": blocks.SNConv(in_channels, num_channels, 1, bias=True)})

        self.blocks = nn.ModuleList()
        num_down_blocks = min(int(math.log(image_size, 2)) - 2, dis_num_blocks)
        cin = num_channels
        cout = cin
        for i in range(1, num_down_blocks):
            cout = min(cin * 2, max_num_channels)
            if i == dis_num_blocks - 1:
                cout = self.out_channels
            self.blocks.append(blocks.PlainResBlock(cin, cout, downsample=True))
            cin = cout
        for i in range(num_down_blocks, dis_num_blocks):
            if i == dis_num_blocks - 1:
                cout = self.out_channels
            self.blocks.append(blocks.PlainResBlock(cin, cout, downsample=False))

        self.linear = blocks.SNLinear(self.out_channels, 1)
        self.embed = SNEmbedding(num_labels, self.out_channels)
        self.finetuning = False
        # The reference's step (runners/holycow.py:239-247) computes discriminator weight gradients during
        # `loss_G.backward()` and then throws them away with `optimizer_D.zero_grad()`.  A runner that follows that
        # protocol may set this flag to run the fake-for-G pass on detached weights: identical losses and identical
        # surviving gradients, one third fewer discriminator weight-gradient GEMMs.
        self.skip_discarded_wgrad = False

    def pass_inputs(self, input, embed=None, detach_params=False):
        """input: (B,3,S,S) NCHW image.  Returns (score (B,), [7 feature maps])  — reference :90-108.
        `detach_params`: run this pass on detached weights (no weight gradients; see `skip_discarded_wgrad`)."""
        x = input.contiguous()
        # one batched power iteration for the 20 MMA convs, the two Cin=3 stem convs and the final linear layer
        blocks.spectral_sigmas(self._tensor_core_convs() + [self.down_block.slot(0), self.skip.slot(0), self.linear])
        w0, s0, b0 = self.down_block.slot(0).operands_edge(detach_params)
        h = ops.conv_c3(x, w0, s0, b0, relu=True, round_out=True)
        w2, s2, b2, e2 = self.down_block.slot(2).operands(detach_params)
        h2 = ops.conv2d(h, w2, s2, bias=b2, ksize=3, **e2)
        # skip: AvgPool2(conv1x1(x)) == conv1x1(AvgPool2(x)); the 1x1 weights ride the centre tap of the 3x3 stem kernel
        ws, ss, bs = self.skip.slot(0).operands_edge(detach_params)
        xs = torch.nn.functional.avg_pool2d(x, 2)
        s = ops.conv_c3(xs, torch.nn.functional.pad(ws, (1, 1, 1, 1)), ss, bs)
        out = ops.avgpool2(h2, s)

        specs = None if os.environ.get('B200LP_NO_DISC_NODE') else self._block_specs(detach_params)
        if specs is not None:
            # the whole block chain as one autograd node with a hand-scheduled backward (ops.DiscBlocksFn)
            feats, out = ops.disc_blocks(out, specs)
        else:
            feats = []
            for block in self.blocks:
                r = ops.relu_round(out)        # the reference's in-place ReLU: this is also what `feats` holds
                feats.append(r)
                out = block(r, detach_params)
        feats.append(out)                  # the last feature stays pre-ReLU (reference :100 is out of place)
        wl, sl, bl = self.linear.operands_edge(detach_params)
        # relu -> spatial sum -> SN-linear + projection onto the label embedding as one kernel (+ two backward)
        score = ops.disc_head(out, embed, wl, sl, bl)
        return score, [f.permute(0, 3, 1, 2) for f in feats]

    def _block_specs(self, detach_params):
        """Operands of every block for ops.DiscBlocksFn, or None when a layer has no batched spectral-norm result (then
        1/sigma carries an autograd edge and the per-layer nodes are used)."""
        for block in self.blocks:
            if any(conv._pre is None or conv._pre[1] is None for conv in block.tensor_core_convs()):
                return None
        specs = []
        for block in self.blocks:
            sp = dict(down=block.downsample, sk=None)
            for key, conv in (("c0", block.block.slot(2)), ("c1", block.block.slot(5)),
                              ("sk", block.skip.slot(0) if block.skip is not None else None)):
                if conv is not None:
                    w, s, b, e = conv.operands(detach_params)
                    sp[key] = (w, s, b, e["cache"], e["sn"])
            specs.append(sp)
        return specs

    def _tensor_core_convs(self):
        """The spectral-normalised convs that run on the tensor cores (everything except the two Cin=3 stem convs)."""
        convs = [self.down_block.slot(2)]
        for block in self.blocks:
            convs += block.tensor_core_convs()
        return convs

    def enable_finetuning(self, data_dict=None):
        """Reference :110-136: the embedding matrix W is replaced by one row initialised from `embeds`."""
        some_parameter = next(iter(self.parameters()))
        if data_dict is None:
            data_dict = {'embeds': torch.rand(1, self.out_channels).to(some_parameter)}
        with torch.no_grad():
            if self.finetuning:
                self.embed.weight_orig.copy_(data_dict['embeds'])
            else:
                new_embed = SNEmbedding(1, self.out_channels, eps=1e-12).to(some_parameter)
                new_embed.weight_orig.copy_(data_dict['embeds'])
                self.embed = new_embed
                self.finetuning = True

    def forward(self, data_dict):
        b200lp_lib.require_device()
        fake_rgbs = data_dict['fake_rgbs']
        target_rgbs = data_dict['target_rgbs']
        label = data_dict['label']
        if len(fake_rgbs.shape) > 4:
            fake_rgbs = fake_rgbs[:, 0]
        if len(target_rgbs.shape) > 4:
            target_rgbs = target_rgbs[:, 0]

        embed = None
        if hasattr(self, 'embed'):
            embed = self.embed(label)

        if self.skip_discarded_wgrad and embed is not None:
            fake_score_G, fake_features = self.pass_inputs(fake_rgbs, embed.detach(), detach_params=True)
        else:
            fake_score_G, fake_features = self.pass_inputs(fake_rgbs, embed)
        fake_score_D, _ = self.pass_inputs(fake_rgbs.detach(), embed.detach())
        real_score, real_features = self.pass_inputs(target_rgbs, embed)

        data_dict['fake_features'] = fake_features
        data_dict['real_features'] = real_features
        data_dict['real_embedding'] = embed
        data_dict['fake_score_G'] = fake_score_G
        data_dict['fake_score_D'] = fake_score_D
        data_dict['real_score'] = real_score
