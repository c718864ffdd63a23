"""Synthetic dataset plugin with the output contract of the reference's
`dataloaders/voxceleb2_segmentation_nolandmarks.py:182-248` (+ default_collate):

    data_dict  : enc_rgbs (K,3,S,S) U[0,1), pose_input_rgbs (1,3,S,S), target_rgbs (1,3,S,S) = image * mask
    target_dict: real_segm (1,3,S,S) in {0,1} (centred disc, r = 0.4 S), label int64

There is no dataset in the sandbox (no network); benchmarks and tests use this plugin (`--dataloader synthetic`).
Sets `args.num_labels` like the reference's dataset code does (dataloaders/common/voxceleb.py:93,107).
"""
import torch


class Dataset(torch.utils.data.Dataset):
    @staticmethod
    def get_args(parser):
        parser.add('--synthetic_num_samples', type=int, default=32)
        parser.add('--synthetic_num_identities', type=int, default=16)
        parser.add('--n_frames_for_encoder', type=int, default=8)
        return parser

    @staticmethod
    def get_dataset(args, part):
        finetune = bool(getattr(args, 'finetune', False))
        num_ids = 1 if finetune else args.synthetic_num_identities
        args.num_labels = num_ids
        k = 1 if finetune else args.n_frames_for_encoder
        return Dataset(args.synthetic_num_samples, args.image_size, k, num_ids, seed=getattr(args, 'random_seed', 123))

    def __init__(self, num_samples, image_size, n_frames_for_encoder, num_identities, seed=123):
        self.num_samples, self.s, self.k, self.num_ids, self.seed = \
            num_samples, image_size, n_frames_for_encoder, num_identities, seed
        s = image_size
        yy, xx = torch.meshgrid(torch.arange(s, dtype=torch.float32), torch.arange(s, dtype=torch.float32), indexing='ij')
        self.mask = (((yy - (s - 1) / 2) ** 2 + (xx - (s - 1) / 2) ** 2) <= (0.4 * s) ** 2).float()

    def __len__(self):
        return self.num_samples

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 100003 + index)
        s = self.s
        image = torch.rand(1, 3, s, s, generator=g)
        data_dict = {
            'enc_rgbs': torch.rand(self.k, 3, s, s, generator=g),
            'pose_input_rgbs': torch.rand(1, 3, s, s, generator=g),
            'target_rgbs': image * self.mask,
        }
        target_dict = {
            'real_segm': self.mask[None, None].expand(1, 3, s, s).contiguous(),
            'label': index % self.num_ids,
        }
        return data_dict, target_dict
