"""Data parallelism over towers: towers are independent graphs, so the batch is
sharded per rank with no data-path collective; the only exchange per training step is ONE
all-reduce of the flat gradient buffer (+ the two loss/accuracy scalars) over NCCL / NVLink.
The reference has no distributed code at all (single process, /root/reference/src/main.py:92-98).
"""
import numpy as np
import torch
import torch.distributed as dist

FLOP_PER_EDGE = 1035600    # forward, reference formulation
FLOP_PER_NODE = 421400


def tower_cost(n_nodes, n_edges):
    return np.asarray(n_edges, dtype=np.float64) * FLOP_PER_EDGE + np.asarray(n_nodes, dtype=np.float64) * FLOP_PER_NODE


def estimate_edges(sizes, fully_connected, mean_degree=3.7):
    sizes = np.asarray(sizes, dtype=np.float64)
    return sizes * (sizes - 1) if fully_connected else np.minimum(sizes * mean_degree, sizes * (sizes - 1))


def shard_towers(sizes, world_size, fully_connected=False):
    """Longest-processing-time assignment of towers to ranks, balanced on the FLOP model.
    Deterministic (stable sort, ties to the lowest rank).  Returns a list of sorted index arrays."""
    sizes = np.asarray(sizes)
    cost = tower_cost(sizes, estimate_edges(sizes, fully_connected))
    order = np.argsort(-cost, kind='stable')
    loads = np.zeros(world_size)
    counts = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(sizes), dtype=np.int64)
    # equal-cost runs are dealt round-robin in O(n); distinct costs go to the least-loaded rank
    for i in order:
        r = int(np.argmin(loads))
        owner[i] = r
        loads[r] += cost[i]
        counts[r] += 1
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world_size)]


def shard_contiguous(n_items, world_size, rank):
    """Equal contiguous split (used when every tower has the same shape)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


class GradientAllReduce:
    """One flat-buffer all-reduce per step.  Works on any backend (nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def broadcast_(self, flat, src=0):
        if self.world > 1:
            dist.broadcast(flat, src, group=self.group)
        return flat

    def allreduce_(self, grads_flat, stats=None):
        if self.world > 1:
            dist.all_reduce(grads_flat, op=dist.ReduceOp.SUM, group=self.group)
            if stats is not None:
                dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return grads_flat, stats


class DataParallelTrainer:
    """engine: spwgnn_b200.engine.Engine on this rank's GPU."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.comm = GradientAllReduce(group)
        self.comm.broadcast_(engine.params.flat, 0)

    def step(self, batch, target, global_count, lr=5e-4):
        """Local forward/backward with the loss normalised by the GLOBAL block count, so the summed
        gradient equals the single-GPU gradient of the whole batch; then all-reduce and Adam."""
        eng = self.engine
        stats = eng.loss_and_grads(batch, target, count=global_count)
        self.comm.allreduce_(eng.grads.flat, stats)
        eng.adam_step(lr=lr)
        return stats
