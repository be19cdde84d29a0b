"""GPU parity of the aggregation pass (al_aggregate) against the reference's
golden outputs (stage-injected with the reference's own logRho and canonical
edge order) and against the CPU oracle on larger seeded inputs."""
import numpy as np
import pytest
import torch

from tests.conftest import golden_names

pytestmark = pytest.mark.gpu


def _native():
    from astrolink_b200 import _native
    _native.require_cuda()
    return _native


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check(nat, lr, ordered_pairs, ref_ordering, ref_groups, ref_prom):
    o, g, p = nat.aggregate(_cu(lr), _cu(ordered_pairs.view(np.int32)))
    torch.cuda.synchronize()
    assert (nat.u32_numpy(o) == ref_ordering).all(), "ordering must be bit-exact"
    gg = nat.u32_numpy(g)
    assert gg.shape == ref_groups.shape and (gg == ref_groups).all(), "groups must be bit-exact"
    pp = p.cpu().numpy()
    tol = 1e-6 if lr.dtype == np.float32 else 1e-14
    assert pp.shape == ref_prom.shape
    if pp.size:
        assert np.max(np.abs(pp.astype(np.float64) - ref_prom.astype(np.float64))) <= tol


@pytest.mark.parametrize("name", golden_names())
def test_aggregate_bit_exact_vs_reference_golden(name, golden):
    nat = _native()
    g = golden(name)
    _check(nat, g["logRho"], g["pairs"][g["order"]], g["ordering"], g["groups"], g["prominences"])


@pytest.mark.parametrize("n,d,dtype,L,seed", [(20000, 3, np.float32, 7, 1), (50000, 6, np.float32, 7, 2), (30000, 2, np.float64, 9, 3),
                                              (200000, 3, np.float32, 7, 4), (8000, 2, np.float32, 2, 5), (8000, 2, np.float64, 3, 6)])
def test_aggregate_vs_oracle_seeded(n, d, dtype, L, seed):
    from oracle import oracle
    nat = _native()
    rng = np.random.default_rng(seed)
    P = np.concatenate([rng.normal(0, 1, (n // 2, d)), rng.normal(2.5, 0.3, (n // 4, d)), rng.uniform(-4, 6, (n - n // 2 - n // 4, d))]).astype(dtype)
    r = oracle.hot_path(P, k_den=20, k_link=L, adaptive=0)
    _check(nat, r["logRho"], r["ordered_pairs"], r["ordering"], r["groups"], r["prominences"])


@pytest.mark.parametrize("seed", range(6))
def test_aggregate_quantised_density_ties(seed):
    """Heavy exact ties in logRho (float32 densities on a coarse grid)."""
    from oracle import oracle
    nat = _native()
    rng = np.random.default_rng(100 + seed)
    n, L = 30000, 6
    P = rng.normal(0, 1, (n, 3)).astype(np.float32)
    d2, idx = oracle.knn(P, 10)
    lr = (np.round(oracle.normalise(oracle.logrho(d2, 10, 3)) * 64) / 64).astype(np.float32)
    pairs, edges = oracle.make_graph(lr, np.ascontiguousarray(idx[:, :L]))
    op = pairs[oracle.sort_edges(edges)]
    o, g, p = oracle.aggregate(lr, op)
    _check(nat, lr, op, o, g, p)
