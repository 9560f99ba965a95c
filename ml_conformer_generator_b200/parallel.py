"""Multi-GPU host logic.  Samples are independent through the whole reverse loop and the GCN, so a batch
is sharded across ranks with NO collective inside the loop; the only exchange is one gather of the results
(coordinates, atom classes, bond matrices).  One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests of this logic) is the plumbing."""
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def sample_cost(n_nodes: np.ndarray) -> np.ndarray:
    """Relative cost of a sample: edge work n(n-1) dominates (F_alg), plus the per-node term."""
    n = np.asarray(n_nodes, dtype=np.float64)
    return 9_593_640.0 * n * (n - 1) + 38_122_560.0 * n


def shard_indices(n_nodes: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Deterministic, cost-balanced partition of sample ids over ranks (longest-processing-time greedy).  Each rank's
    list is sorted, so a rank's samples keep their global order; per-sample RNG is keyed by global id, hence any
    partition reproduces the single-GPU result."""
    n_nodes = np.asarray(n_nodes)
    cost = sample_cost(n_nodes)
    order = np.argsort(-cost, kind="stable")
    loads = np.zeros(world_size)
    buckets: List[List[int]] = [[] for _ in range(world_size)]
    for idx in order:
        r = int(np.argmin(loads))
        buckets[r].append(int(idx))
        loads[r] += cost[idx]
    return [np.sort(np.asarray(b, dtype=np.int64)) for b in buckets]


def contiguous_runs(ids: np.ndarray) -> List[Tuple[int, int]]:
    """[(start, length)] runs of consecutive global ids (the device RNG takes one sample_offset per launch)."""
    runs, start, prev = [], None, None
    for i in ids.tolist():
        if start is None:
            start = prev = i
        elif i == prev + 1:
            prev = i
        else:
            runs.append((start, prev - start + 1))
            start = prev = i
    if start is not None:
        runs.append((start, prev - start + 1))
    return runs


def gather_results(local: Sequence[torch.Tensor], local_ids: np.ndarray, total: int, group=None) -> List[torch.Tensor]:
    """The single collective of the path: all-gather every rank's (ids, tensors...) and scatter them back into global
    sample order.  Shards may differ in size: they are padded to the largest shard for the fixed-size all_gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        outs = []
        for t in local:
            full = torch.zeros((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            full[torch.as_tensor(local_ids, device=t.device)] = t
            outs.append(full)
        return outs
    dev = local[0].device
    n_local = torch.tensor([len(local_ids)], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    cap = int(max(int(s) for s in sizes))
    ids_pad = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    ids_pad[: len(local_ids)] = torch.as_tensor(local_ids, device=dev)
    all_ids = [torch.empty_like(ids_pad) for _ in range(world)]
    dist.all_gather(all_ids, ids_pad, group=group)
    outs = []
    for t in local:
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.zeros((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        for r in range(world):
            k = int(sizes[r])
            full[all_ids[r][:k]] = parts[r][:k]
        outs.append(full)
    return outs


def generate_sharded(run_shard: Callable[[np.ndarray, np.ndarray, int], Sequence[torch.Tensor]], n_nodes: Sequence[int],
                     group=None) -> List[torch.Tensor]:
    """Shard `n_nodes` over the ranks of `group`, call `run_shard(ids, n_nodes[ids], sample_offset)` for every contiguous
    run of this rank's ids, gather.  `run_shard` is the per-GPU hot path (e.g. Engine.generate_host)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_nodes = np.asarray(n_nodes)
    mine = shard_indices(n_nodes, world)[rank]
    pieces: List[Sequence[torch.Tensor]] = []
    for start, length in contiguous_runs(mine):
        ids = np.arange(start, start + length)
        pieces.append(run_shard(ids, n_nodes[ids], start))
    local = [torch.cat([p[k] for p in pieces], dim=0) for k in range(len(pieces[0]))] if pieces else []
    return gather_results(local, mine, len(n_nodes), group)
