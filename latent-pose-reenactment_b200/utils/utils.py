"""Config / plugin loading, checkpoint I/O, meters — the boundary glue of the reference's `utils/utils.py`
(get_args_and_modules :42-164, save_model :251-295, load_model_from_checkpoint :298-398, Meter :196-248),
re-implemented so that the reference's `train.py` / `drive.py` control flow and checkpoint files carry over:

  * argument precedence: argparse defaults < checkpoint args < configs/<name>.yaml < command line;
  * plugins resolved by name with importlib (`generators.<name>`, `criterions.<name>`, ...), each contributing flags
    through `Wrapper.get_args(parser)`;
  * checkpoints are one .pth: {embedder, generator, discriminator, optimizer_G, optimizer_D, running_averages, args}.
"""
import importlib
import logging
import os
import random
import time
from argparse import Namespace
from collections import defaultdict
from pathlib import Path

import torch
import yaml

CONFIG_DIRS = [Path('configs'), Path(__file__).resolve().parent.parent / 'configs']


def setup(args):
    logger = logging.getLogger('utils.setup')
    try:
        import cv2
        cv2.setNumThreads(0)
        cv2.ocl.setUseOpenCL(False)
    except Exception:
        pass
    torch.set_num_threads(1)
    os.environ['OMP_NUM_THREADS'] = '1'
    if args.random_seed is None:
        args.random_seed = int(time.time() * 2)
    logger.info(f"Random Seed: {args.random_seed}")
    random.seed(args.random_seed)
    torch.manual_seed(args.random_seed)
    if str(args.device).startswith('cuda'):
        torch.cuda.manual_seed_all(args.random_seed)


def dict_to_device(d, device):
    for key in d:
        if torch.is_tensor(d[key]):
            d[key] = d[key].to(device, non_blocking=True)


def load_checkpoint_file(path):
    """Checkpoints pickle an argparse.Namespace (and pathlib paths): torch >= 2.6 needs weights_only=False."""
    return torch.load(path, map_location='cpu', weights_only=False)


def load_config_file(config_name):
    logger = logging.getLogger('utils.load_config_file')
    for d in CONFIG_DIRS:
        path = d / f'{config_name}.yaml'
        if path.is_file():
            logger.info(f"Using config {path}")
            with open(path, 'r') as stream:
                return yaml.safe_load(os.path.expandvars(stream.read())) or {}
    raise FileNotFoundError(f"configs/{config_name}.yaml")


def load_module(module_type, module_name):
    return importlib.import_module(f'{module_type}.{module_name}')


def load_wrappers_for_module_list(module_name_list, parent_module):
    names = [c.strip() for c in module_name_list.split(',') if c.strip()]
    return [importlib.import_module(f'{parent_module}.{n}').Wrapper for n in names]


def get_args_and_modules(parser, use_checkpoint_args=True, custom_args={}):
    """Returns (args, default_args, modules dict, checkpoint_object) with the reference's resolution order."""
    logger = logging.getLogger('utils.get_args_and_modules')

    parser.set_defaults(**custom_args)
    args, _ = parser.parse_known_args()
    try:
        config_args = load_config_file(args.config_name) if args.config_name else {}
        if not args.config_name:
            logger.warning("Not using any .yaml config file")
    except FileNotFoundError:
        logger.warning(f"Could not load config {args.config_name}")
        config_args = {}

    parser.set_defaults(**config_args)
    parser.set_defaults(**custom_args)
    args, _ = parser.parse_known_args()

    checkpoint_object, checkpoint_args = None, {}
    if use_checkpoint_args and args.checkpoint_path:
        logger.info(f"Loading checkpoint file {args.checkpoint_path}")
        checkpoint_object = load_checkpoint_file(args.checkpoint_path)
        checkpoint_args = vars(checkpoint_object['args'])

    def resolve(final=False):
        parser.set_defaults(**checkpoint_args)
        parser.set_defaults(**config_args)
        parser.set_defaults(**custom_args)
        return parser.parse_args() if final else parser.parse_known_args()[0]

    args = resolve()
    m = {}
    m['generator'] = load_module('generators', args.generator).Wrapper
    m['generator'].get_args(parser)
    m['embedder'] = load_module('embedders', args.embedder).Wrapper
    m['embedder'].get_args(parser)
    m['runner'] = load_module('runners', args.runner)
    m['runner'].get_args(parser)
    m['discriminator'] = load_module('discriminators', args.discriminator).Wrapper
    m['discriminator'].get_args(parser)
    m['criterion_list'] = load_wrappers_for_module_list(args.criterions, 'criterions')
    for crit in m['criterion_list']:
        crit.get_args(parser)
    m['metric_list'] = load_wrappers_for_module_list(args.metrics, 'metrics')
    for metric in m['metric_list']:
        metric.get_args(parser)
    m['dataloader'] = load_module('dataloaders', 'dataloader').Dataloader(args.dataloader)
    m['dataloader'].get_args(parser)

    args = resolve(final=True)
    default_args = parser.parse_args([])
    if not args.experiment_name:
        args.experiment_name = args.config_name
    return args, default_args, m, checkpoint_object


class Meter:
    """Average / last value of named scalars; NaNs are recorded as `last` but excluded from the average."""

    def __init__(self):
        self.sum = defaultdict(float)
        self.num_measurements = defaultdict(int)
        self.last_value = {}

    def add(self, name, value, num_measurements=1):
        assert num_measurements >= 0
        if num_measurements == 0:
            return
        value = float(value)
        if value != value:
            self.sum[name] += 0
            self.num_measurements[name] += 0
        else:
            self.sum[name] += value * num_measurements
            self.num_measurements[name] += num_measurements
        self.last_value[name] = value

    def keys(self):
        return self.sum.keys()

    def get_average(self, name):
        return self.sum[name] / max(1, self.num_measurements[name])

    def get_last(self, name):
        return self.last_value[name]

    def get_num_measurements(self, name):
        return self.num_measurements[name]

    def __iadd__(self, other):
        for name in other.sum:
            self.add(name, other.get_average(name), other.get_num_measurements(name))
            self.last_value[name] = other.last_value[name]
        return self


def save_model(training_module, optimizer_G, optimizer_D, args):
    logger = logging.getLogger('utils.save_model')
    if args.rank != 0:
        return
    training_module = getattr(training_module, 'module', training_module)
    save_dict = {}
    for name in ('embedder', 'generator', 'discriminator'):
        module = getattr(training_module, name)
        if module is not None:
            save_dict[name] = module.state_dict()
    if optimizer_G is not None:
        save_dict['optimizer_G'] = optimizer_G.state_dict()
    if optimizer_D is not None:
        save_dict['optimizer_D'] = optimizer_D.state_dict()
    if training_module.running_averages is not None:
        save_dict['running_averages'] = {n: m.state_dict() for n, m in training_module.running_averages.items()}
    save_dict['args'] = args

    stem = f'{args.iteration:08}'
    save_path = f'{args.experiment_dir}/checkpoints/model_{stem}.pth'
    while os.path.exists(save_path):
        stem += '_0'
        save_path = f'{args.experiment_dir}/checkpoints/model_{stem}.pth'
    os.makedirs(os.path.dirname(save_path), exist_ok=True)
    try:
        logger.info(f"Saving checkpoint at {save_path}")
        torch.save(save_dict, save_path, pickle_protocol=-1)
    except RuntimeError as err:   # disk full: do not leave a truncated file behind
        logger.error(f"Could not write to {save_path}: {err}; removing that file")
        try:
            os.remove(save_path)
        except OSError:
            pass
    return save_path


def load_model_from_checkpoint(checkpoint_object, args=Namespace()):
    """Rebuild embedder / generator / discriminator (+ optimizers) from a checkpoint, entering or keeping
    fine-tuning mode as the reference does (:298-398)."""
    logger = logging.getLogger('utils.load_model_from_checkpoint')
    saved_args = checkpoint_object['args']
    saved_device, saved_args.device = saved_args.device, 'cpu'

    finetune = 'finetune' in args and args.finetune
    already_finetuned = 'finetune' in saved_args and saved_args.finetune
    assert not (already_finetuned and 'finetune' in args and not finetune), \
        "NYI: using fine-tuned checkpoint for meta-learning"

    differing_args = [k for k, v in vars(args).items() if k in saved_args and v != vars(saved_args).get(k)]
    running_averages = checkpoint_object.get('running_averages', {})

    modules = {}
    for module_name in 'embedder', 'generator', 'discriminator':
        module_kind = getattr(args, module_name)
        logger.info(f"Loading {module_name} '{module_kind}'")
        wrapper = load_module(f'{module_name}s', module_kind).Wrapper
        module = wrapper.get_net(args)
        module_old = wrapper.get_net(saved_args)
        if already_finetuned:
            module_old.enable_finetuning()
        module_old.load_state_dict(checkpoint_object[module_name])
        if finetune:
            module.enable_finetuning()
            if not already_finetuned:
                module_old.enable_finetuning()
        if module_name in differing_args:
            logger.warning(f"{module_name} has changed in config, so not loading weights")
        else:
            module.load_state_dict(module_old.state_dict())
        modules[module_name] = module

    if 'inference' in args and args.inference:
        optimizer_G = optimizer_D = None
    else:
        optimizer_D = load_module('discriminators', args.discriminator).Wrapper \
            .get_optimizer(modules['discriminator'], args)
        if 'discriminator' in differing_args or optimizer_D is None or finetune and not already_finetuned:
            logger.warning("Discriminator has changed in config (maybe due to finetuning), so not loading `optimizer_D`")
        else:
            optimizer_D.load_state_dict(checkpoint_object['optimizer_D'])
        runner = load_module('runners', args.runner)
        optimizer_G = runner.get_optimizer(modules['embedder'], modules['generator'], args)
        if 'generator' in differing_args or 'embedder' in differing_args or finetune and not already_finetuned:
            logger.warning("Embedder or generator has changed in config, so not loading `optimizer_G`")
        else:
            optimizer_G.load_state_dict(checkpoint_object['optimizer_G'])

    saved_args.device = saved_device
    return (modules['embedder'], modules['generator'], modules['discriminator'],
            running_averages, saved_args, optimizer_G, optimizer_D)
