"""CPU: multi-GPU host logic (sharding + the single packed result gather) with the gloo backend, world_size 2."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_conformer_generator_b200 import parallel as P

N_MAX = 4


def test_shard_indices_partition_and_balance():
    rng = np.random.RandomState(0)
    n_nodes = rng.randint(15, 40, size=1000)
    for world in (1, 2, 4, 8):
        shards = P.shard_indices(n_nodes, world)
        allids = np.sort(np.concatenate(shards))
        assert np.array_equal(allids, np.arange(1000))
        loads = np.array([P.sample_cost(n_nodes[s]).sum() for s in shards])
        assert loads.max() / loads.mean() < 1.01  # edge work varies 7x between 15 and 39 atoms; shards stay balanced
        assert all(np.all(np.diff(s) > 0) for s in shards)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    # equal-size samples (config C2): still an exact partition with equal shard sizes
    shards = P.shard_indices(np.full(8192, 39), 8)
    assert all(len(s) == 1024 for s in shards)
    # the config-5 sweep: 65 536 samples of 15..39 atoms over 8 ranks, balanced to 0.1 %
    n5 = np.random.RandomState(5).randint(15, 40, 65536)
    loads = np.array([P.sample_cost(n5[s]).sum() for s in P.shard_indices(n5, 8)])
    assert loads.max() / loads.mean() < 1.001


def test_pack_unpack_round_trip():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 39, 3, generator=g)
    cls = torch.randint(-1, 7, (5, 39), generator=g, dtype=torch.int32)
    bonds = torch.randint(0, 5, (5, 42, 42), generator=g).to(torch.int8)
    buf = P.pack_results(x, cls, bonds)
    assert buf.shape == (5, P.result_bytes(39)) and buf.dtype == torch.uint8
    x2, c2, b2 = P.unpack_results(buf, 39)
    assert torch.equal(x, x2) and torch.equal(cls, c2) and torch.equal(bonds, b2)


def _fake_hot_path(ids, n_nodes):
    """Stand-in for the per-GPU path: a pure function of (global id, atom count), like the keyed device RNG."""
    ids_t = torch.as_tensor(ids, dtype=torch.float32)
    x = ids_t.view(-1, 1, 1) * torch.ones(len(ids), N_MAX, 3) + torch.as_tensor(n_nodes, dtype=torch.float32).view(-1, 1, 1)
    cls = (torch.as_tensor(ids).view(-1, 1) % 7).repeat(1, N_MAX).to(torch.int32)
    bonds = (torch.as_tensor(ids).view(-1, 1, 1) % 5).repeat(1, 42, 42).to(torch.int8)
    return x, cls, bonds


def _worker(rank, world, port, n_nodes, tmp, max_batch):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {}
    outs = P.generate_sharded(_fake_hot_path, n_nodes, N_MAX, max_batch=max_batch, stats=stats)
    torch.save((outs, stats), tmp + ".%d" % rank)
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, n_nodes, port, max_batch=8192):
    single = P.generate_sharded(_fake_hot_path, n_nodes, N_MAX)  # world size 1, no process group
    out_file = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, n_nodes, out_file, max_batch), nprocs=2, join=True)
    res = [torch.load(out_file + ".%d" % r) for r in range(2)]
    for outs, _ in res:  # every rank holds the full result, identical to the single-process one
        for a, b in zip(single, outs):
            assert a.dtype == b.dtype and torch.equal(a, b)
    return [st for _, st in res]


def test_sharded_generation_matches_single_process(tmp_path):
    rng = np.random.RandomState(1)
    n_nodes = rng.randint(15, 40, size=37)
    stats = _run(tmp_path, n_nodes, 29517)
    # one launch sequence per rank (the shard fits one sub-batch), not one per run of consecutive ids
    assert [s["calls"] for s in stats] == [1, 1]
    assert sorted(s["shard"] for s in stats) == [18, 19]


def test_sharded_generation_sub_batches_and_equal_sizes(tmp_path):
    stats = _run(tmp_path, np.full(64, 39), 29518, max_batch=10)
    assert [s["calls"] for s in stats] == [4, 4]  # ceil(32 / 10): bounded by memory, not by id contiguity


def test_sharded_generation_with_an_empty_shard(tmp_path):
    stats = _run(tmp_path, np.array([17]), 29519)
    assert sorted(s["calls"] for s in stats) == [0, 1]
