"""Result saver used by the optional validation phase (observability; out of the hot path)."""
import os

import torch


class Saver:
    def __init__(self, save_dir, save_fn='tensors'):
        self.save_dir = save_dir
        self.save_fn = save_fn
        os.makedirs(save_dir, exist_ok=True)
        self.counter = 0

    def save(self, epoch, data):
        out = {k: v.detach().cpu() for k, v in data.items() if torch.is_tensor(v)}
        torch.save(out, os.path.join(self.save_dir, f'epoch{epoch:04}_{self.counter:06}.pt'))
        self.counter += 1


def make_visual(data_dict, n_samples=2):
    """Side-by-side (target | fake) uint8 HWC image grid for TensorBoard."""
    tgt = data_dict['target_rgbs']
    tgt = tgt[:, 0] if tgt.dim() > 4 else tgt
    fake = data_dict['fake_rgbs']
    rows = [torch.cat((tgt[i], fake[i]), dim=2) for i in range(min(n_samples, len(fake)))]
    grid = torch.cat(rows, dim=1).detach().clamp(0, 1).mul(255).byte().permute(1, 2, 0).cpu().numpy()
    return grid
