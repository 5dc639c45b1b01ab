"""Batch sharding + flat-gradient all-reduce on CPU: gloo, world_size 2.

The compute here is the CPU oracle (tests may use it); what is under test is the host logic the GPU path
uses unchanged: equal shards, one flat FP32 buffer, sum all-reduce, 1/world scaling.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ancde_b200 import dp, synthetic, metamodel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loss(sd, cfg, bt, coeffs, sl):
    from oracle import port
    res = port.dual_ncde_forward(sd, cfg, bt['times'], tuple(c[sl] for c in coeffs), bt['final_index'][sl], 1.0,
                                 stream=cfg['stream'])
    return torch.nn.functional.mse_loss(res['pred'], bt['y'][sl])


def _worker(rank, world, port_no, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import port
    cfg = synthetic.CONFIGS['mujoco']
    torch.manual_seed(2021)
    model, _, _ = metamodel.model_from_config(cfg)
    bucket = dp.FlatGradBucket(model.parameters())
    bt = synthetic.make_batch('mujoco', B=8, seed=3)
    coeffs = port.natural_cubic_spline_coeffs(bt['times'], bt['X'])
    sd = dict(model.named_parameters())
    bucket.zero()
    _loss(sd, cfg, bt, coeffs, dp.shard_slice(8, rank, world)).backward()
    bucket.all_reduce_mean()
    if rank == 0:
        torch.save(bucket.flat.clone(), out)
    dist.destroy_process_group()


def test_two_rank_sharded_gradient_equals_full_batch_gradient(tmp_path):
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sharded = torch.load(out)
    from oracle import port
    cfg = synthetic.CONFIGS['mujoco']
    torch.manual_seed(2021)
    model, _, _ = metamodel.model_from_config(cfg)
    bucket = dp.FlatGradBucket(model.parameters())
    bt = synthetic.make_batch('mujoco', B=8, seed=3)
    coeffs = port.natural_cubic_spline_coeffs(bt['times'], bt['X'])
    _loss(dict(model.named_parameters()), cfg, bt, coeffs, slice(0, 8)).backward()
    assert sharded.shape == bucket.flat.shape
    assert (sharded - bucket.flat).abs().max() <= 1e-5 * bucket.flat.abs().max()


def test_flat_bucket_views_and_shards():
    lin = torch.nn.Linear(3, 2)
    b = dp.FlatGradBucket(lin.parameters())
    assert b.numel == 8 and lin.weight.grad.data_ptr() == b.flat.data_ptr()
    lin(torch.ones(1, 3)).sum().backward()
    assert b.flat.abs().sum() > 0 and torch.equal(lin.weight.grad.view(-1), b.flat[:6])
    b.zero()
    assert b.flat.abs().sum() == 0 and lin.weight.grad.data_ptr() == b.flat.data_ptr()
    assert dp.shard_slice(1024, 3, 8) == slice(384, 512)
    x = torch.arange(10)
    assert dp.shard(x, 1, 2).tolist() == [5, 6, 7, 8, 9]
    with pytest.raises(ValueError):
        dp.FlatGradBucket([])
