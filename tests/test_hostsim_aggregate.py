"""Logic of the parallel aggregation algorithm (astrolink_b200/csrc/aggregate_impl.h)
run by the serial host executor (tests/hostsim) -- no GPU needed -- against the
reference goldens and the oracle.  The GPU executor of the same algorithm is
covered by tests/test_gpu_aggregate.py."""
import numpy as np
import pytest

from oracle import oracle
from tests.conftest import golden_names
from tests.hostsim import hostsim


@pytest.mark.parametrize("name", golden_names())
def test_hostsim_matches_reference_golden(name, golden):
    g = golden(name)
    o, gr, pr = hostsim.aggregate(g["logRho"], g["pairs"][g["order"]])
    assert (o == g["ordering"]).all()
    assert gr.shape == g["groups"].shape and (gr == g["groups"]).all()
    assert np.allclose(pr, g["prominences"], rtol=0, atol=1e-6 if pr.dtype == np.float32 else 1e-14)


def _rand_graph(rng, n, L, with_self):
    W = min(n - 1, 40)
    off = np.argsort(rng.random((n, W)), axis=1)[:, :L - 1 if with_self else L] + 1
    nb = (np.arange(n)[:, None] + off) % n
    if with_self:
        nb = np.concatenate([np.arange(n)[:, None], nb], axis=1)
    pi = rng.permutation(n)
    out = np.empty_like(nb)
    out[pi] = pi[nb]
    return out.astype(np.uint32)


@pytest.mark.parametrize("seed", range(8))
def test_hostsim_matches_oracle_random_graphs(seed):
    rng = np.random.default_rng(seed)
    for trial in range(40):
        n = int(rng.integers(12, 1500))
        L = int(rng.integers(2, 9))
        dtype = np.float32 if trial % 2 else np.float64
        if trial % 3 == 0:      # heavy exact ties
            lr = (rng.integers(0, max(2, n // 10), n) / max(1, n // 10)).astype(dtype)
            knn = _rand_graph(rng, n, L, True)
        elif trial % 3 == 1:    # self missing from the rows
            lr = rng.random(n).astype(dtype)
            knn = _rand_graph(rng, n, L, False)
        else:                   # geometric
            P = np.concatenate([rng.normal(c, 0.3, (n // 3 + 1, 3)) for c in (0, 4, 8)])[:n].astype(dtype)
            d2, idx = oracle.knn(P, max(L, 5))
            lr = oracle.normalise(oracle.logrho(d2, max(L, 5), 3))
            knn = np.ascontiguousarray(idx[:, :L])
        pairs, edges = oracle.make_graph(lr, knn)
        op = pairs[oracle.sort_edges(edges)]
        o1, g1, p1 = oracle.aggregate(lr, op)
        o2, g2, p2 = hostsim.aggregate(lr, op)
        assert (o1 == o2).all()
        assert g1.shape == g2.shape and (g1 == g2).all()
        assert np.allclose(p1, p2, rtol=0, atol=1e-6)


@pytest.mark.parametrize("order", [1, 2])
def test_hostsim_is_independent_of_thread_order(order, monkeypatch):
    """The passes that race benignly on the GPU (pointer jumping, hooking, the Link1/Link2 rewiring) must give the same
    result whichever order their threads run in: the host executor replays every pass descending (1) and in a fixed
    pseudo-random permutation (2) -- two more valid interleavings -- and must still match the oracle."""
    monkeypatch.setenv("HOSTSIM_ORDER", str(order))
    rng = np.random.default_rng(100 + order)
    for trial in range(30):
        n = int(rng.integers(12, 1200))
        L = int(rng.integers(2, 9))
        dtype = np.float32 if trial % 2 else np.float64
        if trial % 2 == 0:
            lr = (rng.integers(0, max(2, n // 10), n) / max(1, n // 10)).astype(dtype)
        else:
            lr = rng.random(n).astype(dtype)
        knn = _rand_graph(rng, n, L, trial % 3 != 1)
        pairs, edges = oracle.make_graph(lr, knn)
        op = pairs[oracle.sort_edges(edges)]
        o1, g1, p1 = oracle.aggregate(lr, op)
        o2, g2, p2 = hostsim.aggregate(lr, op)
        assert (o1 == o2).all()
        assert g1.shape == g2.shape and (g1 == g2).all()
        assert np.allclose(p1, p2, rtol=0, atol=1e-6)
