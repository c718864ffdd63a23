"""Adversarial (hinge) criterion — drop-in for the reference's `criterions/adversarial.py:15-57`.
Operates on three (B,) score vectors: plain tensor arithmetic, no kernel needed."""
import torch
from torch import nn


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--gan_type', type=str, default='gan', help='gan|rgan|ragan')

    @staticmethod
    def get_net(args):
        criterion = Criterion(args.gan_type)
        return criterion.to(args.device)


class Criterion(nn.Module):
    def __init__(self, gan_type):
        super().__init__()
        if gan_type not in ('gan', 'rgan', 'ragan'):
            raise Exception('Incorrect `gan_type` argument')
        self.gan_type = gan_type

    def get_dis_preds(self, real_score, fake_score):
        if self.gan_type == 'gan':
            return real_score, fake_score
        if self.gan_type == 'rgan':
            return real_score - fake_score, fake_score - real_score
        return real_score - fake_score.mean(), fake_score - real_score.mean()

    def forward(self, data_dict):
        fake_score_G = data_dict['fake_score_G']
        fake_score_D = data_dict['fake_score_D']
        real_score = data_dict['real_score']
        if self.gan_type == 'gan':
            # one kernel for both losses (hinge / -mean over the (B,) score vectors), one for their gradients
            from b200lp import ops
            loss_G, loss_D = ops.adversarial_losses(fake_score_G, fake_score_D, real_score)
            return {'adversarial_G': loss_G}, {'adversarial_D': loss_D}

        real_pred, fake_pred_D = self.get_dis_preds(real_score, fake_score_D)
        _, fake_pred_G = self.get_dis_preds(real_score, fake_score_G)

        loss_D = torch.relu(1. - real_pred).mean() + torch.relu(1. + fake_pred_D).mean()
        if self.gan_type == 'gan':
            loss_G = -fake_pred_G.mean()
        else:
            loss_G = torch.relu(1. + real_pred).mean() + torch.relu(1. - fake_pred_G).mean()
        return {'adversarial_G': loss_G}, {'adversarial_D': loss_D}
