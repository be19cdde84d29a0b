"""world_size-2 gloo test of the query-sharding plumbing (shard -> all-gather ->
back to original order) used by the multi-GPU path; the kNN kernel is replaced
by the oracle here because there is no GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as td
    from astrolink_b200 import dist as adist
    from oracle import oracle
    td.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        n, K, L, leaf = 1000, 12, 5, 32
        P = rng.normal(0, 1, (n, 3)).astype(np.float32)
        positions = (n + leaf - 1) // leaf * leaf
        perm = np.full(positions, -1, dtype=np.int64)
        perm[:n] = rng.permutation(n)          # stand-in for the index order
        lo, hi, per = adist.shard_positions(positions, rank, world, leaf)
        d2, idx = oracle.knn(P, K)
        mine = perm[lo:hi]
        knn_s = torch.full((per, L), -1, dtype=torch.int32)
        raw_s = torch.full((per,), float("inf"), dtype=torch.float32)
        valid = mine >= 0
        raw_local = oracle.logrho(d2, K, 3)
        knn_s[:hi - lo][torch.from_numpy(valid)] = torch.from_numpy(idx[mine[valid], :L].view(np.int32))
        raw_s[:hi - lo][torch.from_numpy(valid)] = torch.from_numpy(raw_local[mine[valid]])
        D = adist.TorchDist()
        knn_all = torch.empty((per * world, L), dtype=torch.int32)
        raw_all = torch.empty((per * world,), dtype=torch.float32)
        D.all_gather_into_tensor(knn_all, knn_s)
        D.all_gather_into_tensor(raw_all, raw_s)
        sel = adist.gathered_rows(positions, world, leaf, "cpu")
        kk, rr = knn_all[sel].numpy(), raw_all[sel].numpy()
        out_knn = np.empty((n, L), dtype=np.int32)
        out_raw = np.empty(n, dtype=np.float32)
        ok = perm >= 0
        out_knn[perm[ok]] = kk[ok]
        out_raw[perm[ok]] = rr[ok]
        good = (out_knn.view(np.uint32) == idx[:, :L]).all() and (out_raw == raw_local).all()
        ret[rank] = bool(good)
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_shard_gather_unpermute():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret.get(0) is True and ret.get(1) is True


def test_shard_positions_cover_everything():
    from astrolink_b200 import dist as adist
    for positions, world in [(32, 8), (320, 2), (352, 4), (10016, 3), (3200000, 8)]:
        per = adist.shard_positions(positions, 0, world)[2]
        covered = 0
        for r in range(world):
            lo, hi, p = adist.shard_positions(positions, r, world)
            assert p == per and lo % 32 == 0 and hi % 32 == 0 and hi - lo <= per
            assert lo == min(covered, positions) or lo == covered
            covered = max(covered, hi)
        assert covered == positions


def test_dealt_chunks_cover_every_position_once():
    """deal_chunks/dealt_positions (the interleaved sharding of the peer-memory path): over all ranks the rows
    inside the cloud are exactly the index positions 0..positions-1, each once."""
    from astrolink_b200 import dist as adist
    for positions in (32, 4000 // 32 * 32 + 32, 150016, 10_000_000 // 32 * 32 + 32):
        for world in (1, 2, 3, 4, 8):
            seen = torch.zeros(positions, dtype=torch.int32)
            for r in range(world):
                begin, chunk, stride, rows = adist.deal_chunks(positions, r, world)
                assert chunk % (32 * 4) == 0 and stride == world * chunk and begin == r * chunk
                if rows == 0:
                    assert begin >= positions
                    continue
                assert rows % chunk == 0
                pos = adist.dealt_positions(begin, chunk, stride, rows, "cpu")
                assert int(pos[0]) == begin and int(pos.max()) < positions + chunk
                seen.index_add_(0, pos[pos < positions], torch.ones(int((pos < positions).sum()), dtype=torch.int32))
            assert bool((seen == 1).all()), (positions, world)


def test_weighted_deal_covers_every_position_once():
    """Tail-aware dealing (dist.tail_aware_widths): ranks take different numbers of chunks per round -- rank 0, which runs
    the globally ordered tail, fewer or none -- and the positions are still covered exactly once."""
    from astrolink_b200 import dist as adist
    assert adist.tail_aware_widths(1) == [1] and adist.tail_aware_widths(2) == [3, 5]
    assert adist.tail_aware_widths(3) == [2, 5, 5] and adist.tail_aware_widths(4) == [1, 4, 4, 4]
    assert adist.tail_aware_widths(5) == [0, 1, 1, 1, 1] and adist.tail_aware_widths(8) == [0] + [1] * 7
    for world in range(2, 9):      # rank 0 never takes more than an equal share
        w = adist.tail_aware_widths(world)
        assert len(w) == world and w[0] <= w[1] and len(set(w[1:])) == 1
    for positions in (4000 // 32 * 32 + 32, 150016, 10_000_000 // 32 * 32 + 32):
        for widths in ([3, 5], [0, 1, 1, 1], [0, 1, 1], [2, 0, 1], [0] + [1] * 7, [1, 4, 4, 4]):
            world = len(widths)
            seen = torch.zeros(positions, dtype=torch.int32)
            for r in range(world):
                begin, chunk, stride, rows = adist.deal_chunks(positions, r, world, widths=widths)
                assert stride % (32 * 4) == 0 and chunk % (32 * 4) == 0
                if widths[r] == 0:
                    assert chunk == 0 and rows == 0
                    continue
                if rows == 0:
                    continue
                pos = adist.dealt_positions(begin, chunk, stride, rows, "cpu")
                seen.index_add_(0, pos[pos < positions], torch.ones(int((pos < positions).sum()), dtype=torch.int32))
            assert bool((seen == 1).all()), (positions, widths)
