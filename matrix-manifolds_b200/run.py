#!/usr/bin/env python
"""Experiment driver with the command line and YAML schema of the reference's graphembed/run.py:20-135:

    python run.py --config example_config.yaml [--random_seed 42] [--verbose]
    torchrun --nproc-per-node G run.py --config ...        # pair-sharded over G GPUs

fp64 and everything on the GPU by default (run.py:31-35).  Graph distances come from the BFS kernel
(graphembed.data.load_graph_pdists, same `.cached_pdists/<graph>/cached_pdists.npy` cache), the targets stay
resident in HBM whatever the graph size (the reference moves them to the host for N >= 5000, run.py:45), and the
`Layer_Mean_F1` lazy metric (run.py:83-91) is computed by the GPU FastPrecision (graphembed.pyx, gm_rank_metrics)
right in the validation step instead of on a CPU thread pool.  `graphembed.products.Embedding` configs (Universal
factors with a curvature optimizer) select graphembed.products.TrainingEngine as run.py:74-75 does."""
import argparse
import logging
import os
import random
import sys
from shutil import copyfile

import numpy
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from graphembed.config import parse_config  # noqa: E402
from graphembed.data import GraphDataset, load_graph_pdists  # noqa: E402
from graphembed.train import TrainingEngine  # noqa: E402
from graphembed.utils import Timer, check_mkdir, nnm1d2_to_n  # noqa: E402


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Graph embedding driver.')
    parser.add_argument('--config', type=str, help='The YAML config which sets up this driver.')
    parser.add_argument('--random_seed', type=int, default=42, help='The manual random seed.')
    parser.add_argument('--num_workers', type=int, default=4, help='Accepted for compatibility (lazy CPU metrics).')
    parser.add_argument('--detect_anomaly', action='store_true', help='Enable PyTorch anomaly detection')
    parser.add_argument('--verbose', action='store_true', help='Sets the log level to DEBUG.')
    parser.add_argument('--fp32', action='store_true', help='float32 instead of the float64 default.')
    return parser.parse_args(argv)


def set_seeds(seed):
    torch.manual_seed(seed)
    random.seed(seed)
    numpy.random.seed(seed)


def main(argv=None):
    args = parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO)
    if not torch.cuda.is_available():
        raise SystemExit('graphembed-b200 needs a CUDA device: there is no CPU fallback.')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
        pg = torch.distributed.group.WORLD

    config = parse_config(args.config)
    set_seeds(args.random_seed)
    save_dir = check_mkdir(config['save_dir_root'], increment=(rank == 0)) if rank == 0 else None
    if world > 1:
        box = [save_dir]
        torch.distributed.broadcast_object_list(box, src=0)
        save_dir = box[0]
    if rank == 0:
        copyfile(args.config, os.path.join(save_dir, 'config.yaml'))

    torch.set_default_dtype(torch.float32 if args.fp32 else torch.float64)
    torch.set_default_device(torch.device('cuda', local))
    if args.detect_anomaly:
        torch.autograd.set_detect_anomaly(True)

    gpdists, g = load_graph_pdists(config['input_graph'], cache_dir=config.get('cache_dir') if rank == 0 else None,
                                   device=torch.device('cuda', local))
    n_nodes = nnm1d2_to_n(len(gpdists))
    if 'preprocess' in config:
        gpdists = config['preprocess'](gpdists)
    dataset = GraphDataset(gpdists.to(torch.device('cuda', local)))

    embedding = config['embedding'](n_nodes)
    if not hasattr(embedding, 'manifolds'):
        raise SystemExit('only graphembed.modules.ManifoldEmbedding is on the B200 hot path')
    if world > 1:  # identical replicas
        for t in list(embedding.xs) + list(embedding.curvature_params):
            torch.distributed.broadcast(t.data, src=0)

    optimizers, lr_schedulers = [], []
    if 'embedding_optimizer' in config:
        emb_optim = config['embedding_optimizer'](embedding.xs)
        optimizers.append(emb_optim)
        if 'embedding_lr_scheduler' in config:
            lr_schedulers.append(config['embedding_lr_scheduler'](emb_optim))
    if 'curvature_optimizer' in config:
        curv_optim = config['curvature_optimizer'](embedding.curvature_params)
        optimizers.append(curv_optim)
        if 'curvature_lr_scheduler' in config:
            lr_schedulers.append(config['curvature_lr_scheduler'](curv_optim))

    training_args = dict(embedding=embedding, optimizer=optimizers, lr_scheduler=lr_schedulers,
                         objective_fn=config['objective_fn'], save_dir=save_dir, process_group=pg)
    training_args.update(config['training_params'])
    if 'min_alpha' in training_args or 'max_alpha' in training_args:
        raise SystemExit('deterministic-annealing training (train_da) is outside the B200 hot path')
    engine_cls = TrainingEngine
    if not hasattr(embedding, 'scales'):  # products.Embedding: stabilise every epoch (run.py:74-75)
        from graphembed.products import TrainingEngine as engine_cls
    # gm_rank_metrics sorts one root's N distances inside one SM's shared memory: N <= 32768 with 4-byte keys, 16384
    # with 8-byte keys (include/gm_kernels.h).  A graph it cannot hold trains without the lazy metric (a warning, not
    # an abort at the first validation epoch).
    fp_limit = 32768 if torch.get_default_dtype() == torch.float32 else 16384
    if g is not None and n_nodes > fp_limit and not g.is_directed():
        logging.warning('Layer_Mean_F1 skipped: %d nodes exceed the %d-node limit of the GPU ranking kernel for %s',
                        n_nodes, fp_limit, torch.get_default_dtype())
    if g is not None and n_nodes <= fp_limit and not g.is_directed():
        from concurrent.futures import Future
        from graphembed.pyx import FastPrecision
        with Timer('constructing FastPrecision', loglevel=logging.INFO):
            fp = FastPrecision(g, device=torch.device('cuda', local))

        def layer_mean_f1(mpdists):  # same contract as the reference's pool.submit(...): a future of (means, stds)
            fut = Future()
            fut.set_result(fp.layer_mean_f1_scores(mpdists))
            return fut

        layer_mean_f1.takes_device_tensor = True
        training_args['lazy_metrics'] = {'Layer_Mean_F1': layer_mean_f1}
    engine = engine_cls(**training_args)
    with Timer('training', loglevel=logging.INFO):
        engine(dataset)
    if pg is not None:
        torch.distributed.destroy_process_group()
    return engine


if __name__ == '__main__':
    main()
    sys.exit(0)
