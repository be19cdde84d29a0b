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


DEAL_CHUNK_LEAVES = 1024     # leaves per dealt chunk (32 768 points): small enough to average out regional cost differences,
                             # large enough that a chunk's candidate tiles stay hot in L2


def deal_chunks(positions, rank, world, leaf=32, chunk_leaves=DEAL_CHUNK_LEAVES):
    """Interleaved sharding for the peer-memory exchange: the spatial order is cut into chunks of
    `chunk_leaves` leaves dealt round-robin, rank r owning chunks r, r + world, r + 2 world, ...
    -> (begin, chunk, stride, rows): arguments of al_knn_dealt (index positions) and its number of output rows."""
    n_leaves = positions // leaf
    # small clouds: at least ~8 chunks per rank (chunks are whole CTAs of 4 leaves)
    chunk_leaves = max(4, min(chunk_leaves, ((n_leaves + world * 8 - 1) // (world * 8) + 3) // 4 * 4))
    chunk = chunk_leaves * leaf
    stride = world * chunk
    begin = rank * chunk
    rows = 0 if begin >= positions else (positions - begin + stride - 1) // stride * chunk
    return begin, chunk, stride, rows


def dealt_positions(begin, chunk, stride, rows, device):
    """Index position searched by row i of a dealt launch (may run past the last position in the final chunk)."""
    i = torch.arange(rows, device=device, dtype=torch.int64)
    return begin + (i // chunk) * stride + i % chunk


def gathered_rows(positions, world, leaf, device):
    """Row indices of the concatenated [world * per] gather buffer that hold real
    index positions 0..positions-1, in position order."""
    _, _, per = shard_positions(positions, 0, world, leaf)
    pos = torch.arange(positions, device=device, dtype=torch.int64)
    r = pos // per
    return r * per + (pos - r * per)


class PeerBuffers:
    """Result buffers that live on rank 0's GPU and are mapped into every other rank's address
    space through CUDA IPC (torch's shared-storage mechanism).  Kernels on rank r then store their
    rows straight into rank 0's arrays over NVLink, so the exchange is fused into the compute
    kernels' epilogues and no collective is needed.  Buffers are cached per (name, shape, dtype)."""

    def __init__(self, dist):
        self.dist = dist
        self.cache = {}
        self.epoch = 0

    def next_epoch(self):
        """Collective (every rank, once per exchange).  Exchanges alternate between two sets of buffers: ranks that
        have nothing to do after the exchange start their next run early and may already store its rows while rank 0
        still reads the previous set; they cannot get further ahead than that because the next exchange's barrier
        needs rank 0."""
        self.epoch += 1
        return self.epoch & 1

    def get(self, name, shape, dtype, device):
        """Collective: every rank calls with identical arguments.  Returns a _native.PeerArray backed by rank 0's
        memory (on rank 0 its `.tensor` is a torch view of it)."""
        import torch.distributed as td
        from . import _native as nat
        key = (name, tuple(shape), dtype)
        if key in self.cache:
            return self.cache[key]
        rank = self.dist.get_rank()
        payload = [None]
        arr = None
        with torch.cuda.device(device):
            if rank == 0:
                arr = nat.PeerArray(shape, dtype)
                payload = [arr.handle]
            td.broadcast_object_list(payload, src=0, group=self.dist.group)
            if rank != 0:
                arr = nat.PeerArray(shape, dtype, handle=payload[0])
        self.cache[key] = arr
        return arr


class TorchDist:
    """Thin adaptor over torch.distributed (NCCL on GPUs, gloo in CPU tests).  peer=True (default):
    neighbour lists and densities are written by the kernels directly into rank 0's buffers through
    NVLink peer memory; peer=False: padded shards + NCCL all-gather."""

    def __init__(self, group=None, peer=True):
        import torch.distributed as td
        self.td = td
        self.group = group
        self.peer = PeerBuffers(self) if peer else None

    def get_rank(self):
        return self.td.get_rank(self.group)

    def get_world_size(self):
        return self.td.get_world_size(self.group)

    def all_gather_into_tensor(self, out, inp):
        self.td.all_gather_into_tensor(out, inp, group=self.group)

    def barrier(self):
        self.td.barrier(group=self.group)
