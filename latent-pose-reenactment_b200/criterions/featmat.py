"""Feature-matching criterion — drop-in for the reference's `criterions/featmat.py:15-29`:
mean-L1 between the 7 discriminator feature maps (fake vs real.detach()), averaged over layers, times fm_weight.
Each L1 is one HBM-bound reduction kernel (and one sign kernel in backward) over the NHWC feature memory."""
from torch import nn

from b200lp import ops


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--fm_weight', type=float, default=10.0)

    @staticmethod
    def get_net(args):
        criterion = Criterion(args.fm_weight)
        return criterion.to(args.device)


def _nhwc_memory(t):
    # discriminator features are (B,C,H,W)-shaped views of NHWC memory: permuting back is free
    return t.permute(0, 2, 3, 1) if t.dim() == 4 else t


class Criterion(nn.Module):
    def __init__(self, fm_weight):
        super().__init__()
        self.fm_weight = fm_weight

    def forward(self, data_dict):
        fake_feats = data_dict['fake_features']
        real_feats = data_dict['real_features']
        # mean over the layers of the per-layer mean-L1, times the weight: one node, one accumulator (ops.L1MeanSumFn)
        pairs = [(_nhwc_memory(f), _nhwc_memory(r)) for f, r in zip(fake_feats, real_feats)]
        return {'feature_matching': ops.l1_mean_sum(pairs, self.fm_weight / len(pairs))}
