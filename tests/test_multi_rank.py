"""World-size-2 test of the N > 1 path on CPU (gloo): shard the pair batch, run each shard
independently (here through the oracle - there is no GPU in this suite), reduce the statistics with
the same collective bench.py issues over NCCL, and compare with the single-process answer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openglue_b200.sharding import all_reduce_loss, all_reduce_statistics, match_statistics, shard_pairs, shard_range
from openglue_b200.synthetic import default_config, synthetic_pairs, synthetic_state_dict


def test_shard_range_covers_everything_once():
    for total in (1, 7, 16, 33, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(total, r, world)
                seen += list(range(s, s + c))
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_pairs, out_path):
    from oracle import superglue_oracle as O
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=10)
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(total_pairs, 40, 36, 32, 1, family='planted', seed=5)
    mine = shard_pairs(data, rank, world)
    assert mine['keypoints0'].shape[0] == shard_range(total_pairs, rank, world)[1]
    res = O.run(sd, cfg, mine, 0.2)
    stats = all_reduce_statistics(match_statistics(res['matches0'], res['matching_scores0']))
    gathered = [None] * world
    dist.all_gather_object(gathered, res['matches0'])
    if rank == 0:
        torch.save({'stats': stats, 'matches0': torch.cat(gathered, 0)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    from oracle import superglue_oracle as O
    total = 5                                            # odd on purpose: ranks get 3 and 2 pairs
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=10)
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(total, 40, 36, 32, 1, family='planted', seed=5)
    ref = O.run(sd, cfg, data, 0.2)
    assert torch.equal(got['matches0'], ref['matches0'])            # pairs are independent: sharding changes nothing
    want = all_reduce_statistics(match_statistics(ref['matches0'], ref['matching_scores0']))
    assert got['stats']['pairs'] == total == want['pairs']
    assert abs(got['stats']['matches_per_pair'] - want['matches_per_pair']) < 1e-9
    assert abs(got['stats']['mean_confidence'] - want['mean_confidence']) < 1e-6


def _loss_worker(rank, world, port, total_pairs, out_path):
    """BASELINE.json configs[3] on CPU: labels -> forward -> criterion per shard (oracles), loss reduced over the ranks."""
    from oracle import superglue_oracle as O
    from oracle import gt_matches_oracle as G
    from oracle import loss_oracle as L
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=10)
    sd = synthetic_state_dict(cfg, seed=0)
    data = shard_pairs(synthetic_pairs(total_pairs, 40, 36, 32, 1, family='planted', seed=5), rank, world)
    b = data['keypoints0'].shape[0]
    H = torch.tensor([[0.9, 0.0, 20.0], [0.0, 0.9, 20.0], [0.0, 0.0, 1.0]]).repeat(b, 1, 1)
    g0, g1, _ = G.gt_matches(data['keypoints0'], data['keypoints1'], {'type': ['perspective'] * b, 'H': H})
    scores = O.run(sd, cfg, data, 0.2)['scores']
    out = L.criterion({'gt_matches0': g0, 'gt_matches1': g1}, {'scores': scores})
    red = all_reduce_loss(torch.stack([out['loss'], out['metric_loss']]).double())
    if rank == 0:
        torch.save({'loss': red}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_loss_reduction_matches_single_process(tmp_path):
    from oracle import superglue_oracle as O
    from oracle import gt_matches_oracle as G
    from oracle import loss_oracle as L
    total = 4                                            # equal shards: mean of the rank losses = loss of the whole batch
    out = str(tmp_path / 'loss.pt')
    mp.spawn(_loss_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)['loss']
    cfg = default_config(descriptor_dim=32, num_stages=1, num_iters=10)
    sd = synthetic_state_dict(cfg, seed=0)
    data = synthetic_pairs(total, 40, 36, 32, 1, family='planted', seed=5)
    H = torch.tensor([[0.9, 0.0, 20.0], [0.0, 0.9, 20.0], [0.0, 0.0, 1.0]]).repeat(total, 1, 1)
    g0, g1, _ = G.gt_matches(data['keypoints0'], data['keypoints1'], {'type': ['perspective'] * total, 'H': H})
    want = L.criterion({'gt_matches0': g0, 'gt_matches1': g1}, {'scores': O.run(sd, cfg, data, 0.2)['scores']})
    assert abs(float(got[0]) - float(want['loss'])) < 1e-5 and float(got[1]) == 0.0


def _grad_worker(rank, world, port, out_path):
    """Data-parallel step on CPU: every rank holds the same small module and its own gradients; after
    sharding.all_reduce_gradients every rank holds the mean, bucketed (tiny bucket size -> several buckets, one of them oversized)."""
    from openglue_b200.sharding import all_reduce_gradients
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.Linear(33, 5), torch.nn.Linear(5, 3))
    net[2].bias.requires_grad_(False)                    # a parameter without a gradient is skipped
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g)
    mine = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    nb = all_reduce_gradients(net.parameters(), bucket_bytes=256)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        torch.save({'buckets': nb, 'reduced': [p.grad for p in net.parameters() if p.grad is not None], 'per_rank': gathered}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average(tmp_path):
    out = str(tmp_path / 'grads.pt')
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    assert got['buckets'] >= 3 and len(got['reduced']) == 5
    for i, red in enumerate(got['reduced']):
        want = (got['per_rank'][0][i] + got['per_rank'][1][i]) / 2
        assert torch.allclose(red, want, rtol=0, atol=1e-7)
