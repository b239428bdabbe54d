#!/usr/bin/env python
"""train.py -- same command line as the reference's train.py:12-57, RNN methods only.

    python train.py -d path/to/dataset/ -m RNN --r_t LSTM --r_l 200 --max_length 200 -b 128 ...

Multi-GPU (one process per GPU, gradients all-reduced by libsbr_b200 over NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 train.py ...
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sbr_b200.helpers import command_parser as parse   # noqa: E402
from sbr_b200.helpers.data_handling import DataHandler  # noqa: E402


def training_command_parser(parser):
    parser.add_argument('--tshuffle', help='Shuffle sequences during training.', action='store_true')
    parser.add_argument('--extended_set', help='Use extended training set (contains first half of validation and '
                        'test set).', action='store_true')
    parser.add_argument('-d', dest='dataset', help='Directory name of the dataset.', default='', type=str)
    parser.add_argument('--dir', help='Directory name to save model.', default='', type=str)
    parser.add_argument('--save', choices=['All', 'Best', 'None'], help='Policy for saving models.', default='Best')
    parser.add_argument('--metrics', help='Metrics for validation, comma separated', default='sps', type=str)
    parser.add_argument('--time_based_progress', help='Follow progress based on time rather than iterations.',
                        action='store_true')
    parser.add_argument('--load_last_model', help='Load Last model before starting training.', action='store_true')
    parser.add_argument('--progress', help='Progress intervals', default='2.', type=str)
    parser.add_argument('--mpi', help='Max progress intervals', default=np.inf, type=float)
    parser.add_argument('--max_iter', help='Max number of iterations', default=np.inf, type=float)
    parser.add_argument('--max_time', help='Max training time in seconds', default=np.inf, type=float)
    parser.add_argument('--min_iter', help='Min number of iterations before showing progress', default=0., type=float)


def num(s):
    try:
        return int(s)
    except ValueError:
        return float(s)


def distributed_placement(args):
    """Rank / device / NCCL id of this process when launched by torchrun; single process otherwise."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    device = args.device if args.device is not None else int(os.environ.get('LOCAL_RANK', '0'))
    if world == 1:
        return dict(device=device), None
    from sbr_b200 import _capi
    from sbr_b200.helpers.rendezvous import Control
    ctl = Control()       # control plane (a TCP star next to MASTER_PORT); the data path is the library's NCCL all-reduce
    nccl_id = ctl.broadcast(_capi.nccl_unique_id() if rank == 0 else None)
    return dict(device=device, n_ranks=world, rank=rank, nccl_id=nccl_id, control=ctl), ctl


def main(argv=None):
    args = parse.command_parser(parse.predictor_command_parser, training_command_parser,
                                parse.early_stopping_command_parser, argv=argv)
    placement, dist = distributed_placement(args)
    seed = args.seed if args.seed is not None else (1234 if placement.get('n_ranks', 1) > 1 else None)
    if seed is not None:      # every rank must build the same global batches
        random.seed(seed)
        np.random.seed(seed)
    predictor = parse.get_predictor(args, **placement)
    dataset = DataHandler(dirname=args.dataset, extended_training_set=args.extended_set, shuffle_training=args.tshuffle)
    predictor.prepare_model(dataset)
    result = predictor.train(dataset, save_dir=dataset.dirname + "models/" + args.dir,
                             time_based_progress=args.time_based_progress, progress=num(args.progress),
                             autosave=args.save, max_progress_interval=args.mpi, max_iter=args.max_iter,
                             min_iterations=args.min_iter, max_time=args.max_time,
                             early_stopping=parse.get_early_stopper(args), load_last_model=args.load_last_model,
                             validation_metrics=args.metrics.split(','))
    if dist is not None:
        dist.barrier()
        dist.close()
    return result


if __name__ == '__main__':
    main()
