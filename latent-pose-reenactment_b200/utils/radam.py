"""Rectified Adam with the update rule of the optimizer the reference vendors and installs as `torch.optim.RAdam`
(reference utils/radam.py:29-95, used by configs/finetuning-base.yaml): variance-rectified step when the SMA length
N_sma >= 5, plain momentum-SGD step before that (`degenerated_to_sgd`).  State keys (`step`, `exp_avg`,
`exp_avg_sq`) match, so optimizer state in checkpoints interchanges.  Implemented with multi-tensor (_foreach) ops.
"""
import math

import torch
from torch.optim.optimizer import Optimizer


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        self.degenerated_to_sgd = degenerated_to_sgd
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @staticmethod
    def _step_size(step, beta1, beta2, degenerated_to_sgd):
        beta2_t = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma >= 5:
            rect = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2))
            return n_sma, rect / (1 - beta1 ** step)
        if degenerated_to_sgd:
            return n_sma, 1.0 / (1 - beta1 ** step)
        return n_sma, -1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            by_step = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('RAdam does not support sparse gradients')
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, dtype=torch.float32)
                    state['exp_avg_sq'] = torch.zeros_like(p, dtype=torch.float32)
                state['step'] += 1
                by_step.setdefault(state['step'], []).append(p)
            for step, ps in by_step.items():
                grads = [p.grad.float() for p in ps]
                m = [self.state[p]['exp_avg'] for p in ps]
                v = [self.state[p]['exp_avg_sq'] for p in ps]
                torch._foreach_mul_(v, beta2)
                torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)
                torch._foreach_mul_(m, beta1)
                torch._foreach_add_(m, grads, alpha=1 - beta1)
                n_sma, step_size = self._step_size(step, beta1, beta2, self.degenerated_to_sgd)
                if n_sma >= 5 or step_size > 0:
                    if group['weight_decay'] != 0:
                        torch._foreach_mul_(ps, 1 - group['weight_decay'] * group['lr'])
                if n_sma >= 5:
                    denom = torch._foreach_sqrt(v)
                    torch._foreach_add_(denom, group['eps'])
                    torch._foreach_addcdiv_(ps, m, denom, value=-step_size * group['lr'])
                elif step_size > 0:
                    torch._foreach_add_(ps, m, alpha=-step_size * group['lr'])
        return loss
