"""Entry point: meta-training and fine-tuning — the reference's `train.py` (args :22-92, distributed init :98-118,
model build / checkpoint load :137-173, SIGINT/SIGTERM checkpointing :175-194, fine-tune initialisation :218-279,
epoch loop :284-310) driving the B200-native plugins.

    python train.py --config default --dataloader synthetic                       # 1 GPU
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 train.py --config default --num_gpus 8 --dataloader synthetic
    python train.py --config finetuning-base --checkpoint_path <ckpt> --dataloader synthetic

One process per GPU; rank / world size come from torchrun's environment (or the legacy `--local_rank` flag);
gradients are exchanged with NCCL all-reduce (runners/holycow.py), not apex / horovod.
"""
import os
import sys

os.environ.setdefault('OMP_NUM_THREADS', '1')
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import logging
import signal
from pathlib import Path

import torch

from utils import utils
from utils.argparse_utils import MyArgumentParser
from utils.utils import get_args_and_modules, load_model_from_checkpoint, save_model, setup

logging.basicConfig(level=logging.INFO, stream=sys.stdout,
                    format="PID %(process)d - %(asctime)s - %(levelname)s - %(name)s - %(message)s")
logger = logging.getLogger('train.py')


def build_parser():
    parser = MyArgumentParser(conflict_handler='resolve')
    parser.add('--config_name', type=str, default="")
    for name in ('generator', 'embedder', 'discriminator', 'criterions', 'metrics', 'dataloader', 'runner'):
        parser.add(f'--{name}', type=str, default="", help='')
    parser.add('--args-to-ignore', type=str,
               default="checkpoint,splits_dir,experiments_dir,extension,experiment_name,rank,local_rank,world_size")
    parser.add('--experiments_dir', type=Path, default="data/experiments", help='')
    parser.add('--experiment_name', type=str, default="", help='')
    parser.add('--train_split_path', default="data/splits/train.csv", type=Path)
    parser.add('--val_split_path', default="data/splits/val.csv", type=Path)
    parser.add('--vgg_weights_dir', default="criterions/common/", type=str)
    # Training process
    parser.add('--num_epochs', type=int, default=10 ** 9)
    parser.add('--set_eval_mode_in_train', action='store_bool', default=False)
    parser.add('--set_eval_mode_in_test', action='store_bool', default=True)
    parser.add('--save_frequency', type=int, default=1,
               help="Save checkpoint every X epochs. If 0, save only at the end of training")
    parser.add('--logging', action='store_bool', default=True)
    parser.add('--skip_eval', action='store_bool', default=True)
    parser.add('--profile_flops', action='store_bool', default=False)
    parser.add('--weights_running_average', action='store_bool', default=True)
    parser.add('--finetune', action='store_bool', default=False)
    parser.add('--inference', action='store_bool', default=False)
    # Model
    parser.add('--in_channels', type=int, default=3)
    parser.add('--out_channels', type=int, default=3)
    parser.add('--num_channels', type=int, default=64)
    parser.add('--max_num_channels', type=int, default=512)
    parser.add('--embed_channels', type=int, default=512)
    parser.add('--pose_embedding_size', type=int, default=136)
    parser.add('--image_size', type=int, default=256)
    # Optimizer
    parser.add('--optimizer', default='Adam', type=str, choices=['Adam', 'RAdam'])
    parser.add('--lr_gen', default=5e-5, type=float)
    parser.add('--beta1', default=0.0, type=float, help='beta1 for Adam')
    # Hardware
    parser.add('--device', type=str, default='cuda')
    parser.add('--num_gpus', type=int, default=1, help='processes = GPUs of one node (NCCL all-reduce)')
    parser.add('--rank', type=int, default=0, help='global rank, DO NOT SET')
    parser.add('--local_rank', '--local-rank', type=int, default=0, help='"rank" within a machine, DO NOT SET')
    parser.add('--world_size', type=int, default=1, help='number of devices, DO NOT SET')
    # Misc
    parser.add('--random_seed', type=int, default=123, help='')
    parser.add('--checkpoint_path', type=str, default='')
    parser.add('--saver', type=str, default='')
    return parser


def init_distributed(args):
    if args.num_gpus == 1:
        args.rank = args.local_rank = 0
        args.world_size = 1
        if str(args.device).startswith('cuda'):
            torch.cuda.set_device(0)
        return
    if args.num_gpus > 8:
        raise NotImplementedError("more than one node (the reference's horovod path) is out of scope")
    args.local_rank = int(os.environ.get('LOCAL_RANK', args.local_rank))
    args.rank = int(os.environ.get('RANK', args.local_rank))
    args.world_size = args.num_gpus
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if str(args.device).startswith('cuda'):
        torch.cuda.set_device(args.local_rank)
        args.device = f'cuda:{args.local_rank}'
        backend = 'nccl'
    else:
        backend = 'gloo'
    torch.distributed.init_process_group(backend=backend, init_method='env://', rank=args.rank,
                                         world_size=args.world_size)


def main():
    parser = build_parser()
    args, default_args, m, checkpoint_object = get_args_and_modules(parser, use_checkpoint_args=True)
    setup(args)
    init_distributed(args)
    logger.info(f"Initialized the process group, my rank is {args.rank}")

    if args.finetune and args.num_gpus > 1:
        if args.local_rank == 0:
            logger.warning("Multi-GPU fine-tuning is NYI (as in the reference), setting `--num_gpus=1`")
            args.num_gpus = 1
        else:
            sys.exit()

    logger.info(f"Loading dataloader '{args.dataloader}'")
    dataloader_train = m['dataloader'].get_dataloader(args, part='train', phase='train')
    runner = m['runner']

    if args.checkpoint_path != "":
        if checkpoint_object is None:
            raise FileNotFoundError(f"Checkpoint `{args.checkpoint_path}` not found")
        logger.info(f"Starting from checkpoint {args.checkpoint_path}")
        embedder, generator, discriminator, running_averages, saved_args, optimizer_G, optimizer_D = \
            load_model_from_checkpoint(checkpoint_object, args)
    else:
        if args.finetune:
            logger.error("`--finetune` is set, but `--checkpoint_path` isn't. This has to be a mistake.")
        discriminator = m['discriminator'].get_net(args)
        generator = m['generator'].get_net(args)
        embedder = m['embedder'].get_net(args)
        running_averages = {}
        optimizer_G = runner.get_optimizer(embedder, generator, args)
        optimizer_D = m['discriminator'].get_optimizer(discriminator, args)

    criterion_list = [crit.get_net(args) for crit in m['criterion_list']]
    if not args.weights_running_average:
        running_averages = None

    writer = None
    if args.logging and args.rank == 0:
        from utils.tensorboard_logging import setup_logging
        args.experiment_dir, writer = setup_logging(args, default_args, args.args_to_ignore.split(','))
        args.experiment_dir = Path(args.experiment_dir)
        metric_list = [metric.get_net(args) for metric in m['metric_list']]
    else:
        metric_list = []
        if args.rank == 0:
            args.experiment_dir = Path(args.experiments_dir) / (args.experiment_name or 'experiment')

    training_module = runner.TrainingModule(embedder, generator, discriminator, criterion_list, metric_list,
                                            running_averages)
    training_module.broadcast_parameters()

    # If someone tries to terminate the program, save the weights first (rank 0, parent process only)
    state = {'saved': False}
    if args.rank == 0:
        parent_pid = os.getpid()

        def save_last_model_and_exit(_signum, _frame):
            if state['saved'] or os.getpid() != parent_pid:
                return
            state['saved'] = True
            logger.info("Interrupted, saving the current model")
            save_model(training_module, optimizer_G, optimizer_D, args)
            if writer is not None:
                writer.close()
            sys.exit()

        signal.signal(signal.SIGINT, save_last_model_and_exit)
        signal.signal(signal.SIGTERM, save_last_model_and_exit)

    saver = None
    if args.saver and args.rank == 0:
        from utils.visualize import Saver
        saver = Saver(save_dir=f'{args.experiment_dir}/validation_results/', save_fn=args.saver)

    if args.finetune:
        logger.info(f"For fine-tuning, computing an averaged identity embedding from "
                    f"{len(dataloader_train.dataset)} frames")
        training_module.eval()
        identity_embeddings = []
        with torch.no_grad():
            for data_dict, _ in dataloader_train:
                embedder_avg = training_module.running_averages.get('embedder', training_module.embedder) \
                    if training_module.running_averages else training_module.embedder
                utils.dict_to_device(data_dict, args.device)
                embedder_avg.get_identity_embedding(data_dict)
                identity_embeddings.append(data_dict['embeds_elemwise'].view(-1, args.embed_channels))
            identity_embedding = torch.cat(identity_embeddings).mean(0)
        data_dict = {'embeds': identity_embedding[None]}
        training_module.generator.enable_finetuning(data_dict)
        training_module.discriminator.enable_finetuning(data_dict)
        training_module.embedder.enable_finetuning()
        # The embedder is not in optimizer_G while fine-tuning (runners/holycow.py:35-37): the reference still back-propagates
        # into it and never reads those gradients.  Freezing it leaves every loss and every G / D update unchanged and lets
        # the pose encoder run its forward-only kernel schedule (embedders/mobilenet_native.py).
        training_module.embedder.requires_grad_(False)
        if args.weights_running_average:
            for name in ('generator', 'embedder'):
                if name in training_module.running_averages:
                    training_module.running_averages[name].enable_finetuning(
                        {'embeds': identity_embedding[None].clone()})
        else:
            training_module.initialize_running_averages(None)
        optimizer_G = runner.get_optimizer(training_module.embedder, training_module.generator, args)
        optimizer_D = m['discriminator'].get_optimizer(discriminator, args)

    logger.info("Entering training loop")
    for epoch in range(0, args.num_epochs):
        training_module.train(not args.set_eval_mode_in_train)
        torch.set_grad_enabled(True)
        runner.run_epoch(dataloader_train, training_module, optimizer_G, optimizer_D, epoch, args,
                         phase='train', writer=writer, saver=saver)
        if not args.skip_eval:
            raise NotImplementedError("NYI: validation (as in the reference)")
        if args.rank == 0:
            will_save = epoch == args.num_epochs - 1
            if args.save_frequency != 0:
                will_save |= epoch % args.save_frequency == 0
            if will_save:
                save_model(training_module, optimizer_G, optimizer_D, args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
