"""Dataset-plugin front end — the reference's `dataloaders/dataloader.py:9-49` contract: a dataset plugin module
defines class `Dataset` with `get_args(parser)` and either `get_dataloader(args, part)` or `get_dataset(args, part)`;
in the latter case samples are strided across ranks (`rank, rank+world, ...`) and batched `batch_size // num_gpus`
per process."""
import logging

import torch
from torch.utils.data import DataLoader

from utils.utils import load_module

logger = logging.getLogger('dataloaders.dataloader')


class Dataloader:
    def __init__(self, dataset_name):
        self.dataset = load_module('dataloaders', dataset_name).__dict__['Dataset']

    def get_args(self, parser):
        parser.add('--num_workers', type=int, default=4, help='Number of data loading workers.')
        parser.add('--prefetch_size', type=int, default=16, help='Prefetch queue size')
        parser.add('--batch_size', type=int, default=64, help='Batch size')
        return self.dataset.get_args(parser)

    def get_dataloader(self, args, part, phase):
        if hasattr(self.dataset, 'get_dataloader'):
            return self.dataset.get_dataloader(args, part)
        dataset = self.dataset.get_dataset(args, part)
        assert len(dataset) % args.world_size == 0, \
            "`dataset.get_dataset()` was expected to return a dataset equally divisible by `args.world_size`"
        dataset = torch.utils.data.Subset(dataset, range(args.rank, len(dataset), args.world_size))
        logger.info(f"This process will receive a dataset with {len(dataset)} samples")
        if len(dataset) < args.batch_size:
            logger.warning(f"Dataset length is smaller than batch size ({len(dataset)} < {args.batch_size}), "
                           f"reducing the latter to {len(dataset)}")
            args.batch_size = len(dataset)
        workers = args.num_workers
        return DataLoader(dataset, batch_size=args.batch_size // args.num_gpus, num_workers=workers,
                          pin_memory=True, drop_last=(phase == 'train'), shuffle=(part == 'train'),
                          prefetch_factor=max(2, args.prefetch_size // max(workers, 1)) if workers > 0 else None,
                          persistent_workers=workers > 0)
