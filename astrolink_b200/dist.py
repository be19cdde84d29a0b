"""Query sharding helpers for the multi-GPU path.

kNN + raw density shard naturally: queries are independent given a replicated
point set.  Index positions (the spatially sorted order, in whole 32-point
leaves) are split into `world` contiguous blocks of equal padded length `per`;
each rank searches its block, and the k_link neighbour lists and raw densities
are exchanged with one NCCL all-gather each.  Aggregation runs on rank 0.
"""
import torch


def shard_positions(positions, rank, world, leaf=32):
    """-> (lo, hi, per): this rank owns index positions [lo, hi); every rank's
    block is padded to `per` rows for the equal-size all-gather."""
    n_leaves = positions // leaf
    per_leaves = (n_leaves + world - 1) // world
    lo = min(rank * per_leaves, n_leaves) * leaf
    hi = min((rank + 1) * per_leaves, n_leaves) * leaf
    return lo, hi, per_leaves * leaf


def gathered_rows(positions, world, leaf, device):
    """Row indices of the concatenated [world * per] gather buffer that hold real
    index positions 0..positions-1, in position order."""
    _, _, per = shard_positions(positions, 0, world, leaf)
    pos = torch.arange(positions, device=device, dtype=torch.int64)
    r = pos // per
    return r * per + (pos - r * per)


class TorchDist:
    """Thin adaptor over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as td
        self.td = td
        self.group = group

    def get_rank(self):
        return self.td.get_rank(self.group)

    def get_world_size(self):
        return self.td.get_world_size(self.group)

    def all_gather_into_tensor(self, out, inp):
        self.td.all_gather_into_tensor(out, inp, group=self.group)

    def barrier(self):
        self.td.barrier(group=self.group)
