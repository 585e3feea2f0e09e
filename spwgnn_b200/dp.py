"""Data parallelism over towers: towers are independent graphs, so the batch is
sharded per rank with no data-path collective; the only exchange per training step is ONE
all-reduce of the flat gradient buffer, with the loss / accuracy sums riding in its tail, over
NCCL / NVLink.  The reference has no distributed code at all (single process,
/root/reference/src/main.py:92-98).
"""
import numpy as np
import torch
import torch.distributed as dist

from .params import FLAT_SIZE, STATS_TAIL  # noqa: F401

FLOP_PER_EDGE = 1035600    # forward, reference formulation
FLOP_PER_NODE = 421400


def tower_cost(n_nodes, n_edges):
    return np.asarray(n_edges, dtype=np.float64) * FLOP_PER_EDGE + np.asarray(n_nodes, dtype=np.float64) * FLOP_PER_NODE


def estimate_edges(sizes, fully_connected, mean_degree=3.7):
    """Relation count when the real one is not known yet (fully connected: exact)."""
    sizes = np.asarray(sizes, dtype=np.float64)
    return sizes * (sizes - 1) if fully_connected else np.minimum(sizes * mean_degree, sizes * (sizes - 1))


def tower_edge_counts(batch):
    """Real relation count of every tower of a packed batch: the differences of the per-tower prefix sum that
    spw_edges_count leaves on the device (one (T + 1)-int read-back)."""
    if batch.edge_off is None:
        raise ValueError('this batch was not built by spw_edges_count')
    return np.diff(batch.edge_off.cpu().numpy().astype(np.int64))


def shard_towers(sizes, world_size, fully_connected=False, edges=None):
    """Towers -> ranks, balanced on the FLOP model.  Vectorised: towers are sorted by cost (stable) and dealt to the
    ranks in serpentine order (0..W-1, W-1..0, ...), which pairs every expensive tower of one sweep with a cheap one
    of the next; for thousands of towers per rank the loads agree to a fraction of a percent, in O(n log n) numpy
    (a 65 536-tower batch takes a few milliseconds; the round-1 greedy loop took 0.25 s).  `edges`: the real
    relation counts (tower_edge_counts) -- otherwise they are estimated from the sizes.  Deterministic.
    Returns a list of sorted index arrays."""
    sizes = np.asarray(sizes)
    e = estimate_edges(sizes, fully_connected) if edges is None else np.asarray(edges)
    cost = tower_cost(sizes, e)
    order = np.argsort(-cost, kind='stable')
    pos = np.arange(len(sizes), dtype=np.int32)
    sweep, k = pos // world_size, pos % world_size
    owner = np.empty(len(sizes), dtype=np.int8 if world_size < 128 else np.int32)
    owner[order] = np.where(sweep % 2 == 0, k, world_size - 1 - k)
    by_rank = np.argsort(owner, kind='stable')               # stable: ascending tower index inside every rank
    cuts = np.cumsum(np.bincount(owner, minlength=world_size))[:-1]
    return np.split(by_rank, cuts)


def shard_contiguous(n_items, world_size, rank):
    """Equal contiguous split (used when every tower has the same shape)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


class GradientAllReduce:
    """ONE flat-buffer all-reduce per step.  Works on any backend (nccl on GPUs, gloo in CPU tests).
    allreduce_(grads_flat, stats): `stats` (float64[2]: loss sum, correct count) travels as two extra floats behind the
    gradients when the buffer has a tail (engine.grads_buffer), otherwise in a second, tiny all-reduce."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def broadcast_(self, flat, src=0):
        if self.world > 1:
            dist.broadcast(flat, src, group=self.group)
        return flat

    def allreduce_(self, grads_flat, stats=None, buffer=None):
        """buffer: the (FLAT_SIZE + STATS_TAIL,) tensor `grads_flat` is a view of, or None."""
        if self.world > 1:
            if buffer is not None and stats is not None:
                buffer[FLAT_SIZE:FLAT_SIZE + 2] = stats.to(buffer.dtype)
                dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=self.group)
                stats.copy_(buffer[FLAT_SIZE:FLAT_SIZE + 2].to(stats.dtype))
            else:
                dist.all_reduce(grads_flat, op=dist.ReduceOp.SUM, group=self.group)
                if stats is not None:
                    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return grads_flat, stats


class DataParallelTrainer:
    """engine: spwgnn_b200.engine.Engine on this rank's GPU."""

    def __init__(self, engine, group=None, dropout_rate=0.1, seed=0x5EED0001):
        self.engine = engine
        self.comm = GradientAllReduce(group)
        self.comm.broadcast_(engine.params.flat, 0)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dropout_rate = dropout_rate      # Dropout(0.1) on both encodings, train only (Networks.py:77-78)
        self._seed = seed
        self._step = 0

    def dropout_seed(self):
        """A fresh 64-bit seed per step and per rank (shards must not reuse the same mask indices)."""
        x = (self._seed + 0x9E3779B97F4A7C15 * (self._step * 1024 + self.rank + 1)) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 27; x = (x * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)

    def step(self, batch, target, global_count, lr=5e-4, dropout_rate=None):
        """Local forward/backward with the loss normalised by the GLOBAL block count, so the summed
        gradient equals the single-GPU gradient of the whole batch; then one all-reduce and Adam.
        Dropout as in the reference's fit (rate 0.1 unless overridden; 0 switches it off)."""
        eng = self.engine
        rate = self.dropout_rate if dropout_rate is None else dropout_rate
        self._step += 1
        stats = eng.loss_and_grads(batch, target, count=global_count, dropout_rate=rate, dropout_seed=self.dropout_seed())
        self.comm.allreduce_(eng.grads.flat, stats, buffer=eng.grads_buffer)
        eng.adam_step(lr=lr)
        return stats
