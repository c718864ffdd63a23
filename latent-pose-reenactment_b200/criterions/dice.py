"""Segmentation dice criterion — drop-in for the reference's `criterions/dice.py:15-39`.
-log( 2*sum(f*r) / (sum(f^2) + sum(r^2)) ) * dice_weight, `real_segm` (B,3,S,S) broadcasting against `fake_segm`
(B,1,S,S) exactly as in the reference."""
import torch
from torch import nn


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--dice_weight', type=float, default=1)

    @staticmethod
    def get_net(args):
        criterion = Criterion(args.dice_weight)
        return criterion.to(args.device)


class Criterion(nn.Module):
    def __init__(self, dice_weight):
        super().__init__()
        self.dice_weight = dice_weight

    def forward(self, data_dict):
        fake_segm = data_dict['fake_segm']
        real_segm = data_dict['real_segm']
        if len(fake_segm.shape) > 4:
            fake_segm = fake_segm[:, 0]
        if len(real_segm.shape) > 4:
            real_segm = real_segm[:, 0]
        if fake_segm.dim() == 4 and fake_segm.shape[1] == 1 and real_segm.dim() == 4:
            from b200lp import ops      # three batch sums + the gradient as two kernels each way
            return {'segmentation_dice': ops.dice_loss(fake_segm, real_segm, self.dice_weight)}
        numer = (2 * fake_segm * real_segm).sum()
        denom = (fake_segm ** 2).sum() + (real_segm ** 2).sum()
        return {'segmentation_dice': -torch.log(numer / denom) * self.dice_weight}
