"""Experiment directory + optional TensorBoard writer (observability only; out of the hot path)."""
import logging
import os
import time
from pathlib import Path

logger = logging.getLogger('utils.tensorboard_logging')


class NullWriter:
    """Stands in when tensorboard is unavailable; keeps `args.iteration` bookkeeping of the runner working."""

    def add_scalar(self, *a, **k): pass
    def add_image(self, *a, **k): pass
    def add_text(self, *a, **k): pass
    def flush(self): pass
    def close(self): pass


def get_experiment_name(args, default_args, args_to_ignore):
    parts = [str(args.experiment_name)] if getattr(args, 'experiment_name', '') else []
    for key, value in sorted(vars(args).items()):
        if key in args_to_ignore or key in ('experiment_name', 'config_name'):
            continue
        if key in vars(default_args) and vars(default_args)[key] != value and isinstance(value, (int, float, bool)):
            parts.append(f"{key}={value}")
    return ('_'.join(parts) or 'experiment')[:120]


def setup_logging(args, default_args, args_to_ignore):
    """Creates `<experiments_dir>/<name>_<timestamp>/checkpoints` and returns (experiment_dir, writer)."""
    name = get_experiment_name(args, default_args, args_to_ignore)
    experiment_dir = Path(args.experiments_dir) / f"{name}_{time.strftime('%Y-%m-%d_%H-%M-%S')}"
    os.makedirs(experiment_dir / 'checkpoints', exist_ok=True)
    try:
        from torch.utils.tensorboard import SummaryWriter
        writer = SummaryWriter(str(experiment_dir / 'tensorboard'), flush_secs=10)
    except Exception as err:  # tensorboard not installed
        logger.warning(f"TensorBoard unavailable ({err}); logging scalars to stdout only")
        writer = NullWriter()
    return str(experiment_dir), writer
