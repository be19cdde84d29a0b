"""Degenerate inputs: >= k_den coincident points, a zero-variance feature, neighbour rows
that do not contain their own point.

Where the reference DEFINES the result (the raw log-density, plain @njit, IEEE semantics: astrolink.py:293-303) the
expected values are the reference's own outputs (tests/golden/degenerate/, written by oracle/gen_degenerate.py).
Where it does not (fastmath min/max over NaN :305-313, the uninitialised column of a zero-variance feature :239-242,
the row overrun of _make_graph_njit :368-380) the tests pin this package's documented behaviour.
"""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_DIR

ROWS = os.path.join(GOLDEN_DIR, "degenerate", "density_rows.npz")
# rows of the fixture: 0-63 ordinary, 64-67 all zero (core distance 0), 68-71 half zero, 72-75 all equal (knife edge:
# k - sum/core cancels to +-ulp, the sign depends on the summation order -- not compared), 76-79 denormal, 80-83 huge
ZERO_CORE = slice(64, 68)
KNIFE_EDGE = slice(72, 76)


def _check_density(got, ref, rtol):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    keep = np.ones(ref.shape[0], dtype=bool)
    keep[KNIFE_EDGE] = False
    assert np.isnan(got[ZERO_CORE]).all() and np.isnan(ref[ZERO_CORE]).all(), "a zero core distance gives NaN, as in the reference"
    g, r = got[keep], ref[keep]
    assert (np.isnan(g) == np.isnan(r)).all()
    assert (np.isneginf(g) == np.isneginf(r)).all() and (np.isposinf(g) == np.isposinf(r)).all()
    fin = np.isfinite(r)
    assert np.all(np.abs(g[fin] - r[fin]) <= rtol * np.maximum(np.abs(r[fin]), 1e-30))


@pytest.mark.parametrize("tag,rtol", [("f32", 2e-6), ("f64", 1e-13)])
def test_oracle_density_on_degenerate_rows(tag, rtol):
    from oracle import oracle
    z = np.load(ROWS)
    d2 = z[f"d2_{tag}"]
    for d_int in (1, 3, 6):
        _check_density(oracle.logrho(d2, d2.shape[1], d_int), z[f"raw_{tag}_d{d_int}"].astype(d2.dtype), rtol)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,rtol", [("f32", 2e-6), ("f64", 1e-13)])
def test_density_kernel_on_degenerate_rows(tag, rtol):
    from astrolink_b200 import _native as nat
    nat.require_cuda()
    z = np.load(ROWS)
    d2 = torch.from_numpy(z[f"d2_{tag}"]).cuda()
    K = d2.shape[1]
    for d_int in (1, 3, 6):
        raw = nat.logrho(d2, None, K, d_int)
        _check_density(raw.cpu().numpy(), z[f"raw_{tag}_d{d_int}"].astype(z[f"d2_{tag}"].dtype), rtol)
    # weighted (a5): the weights are gathered through an identity index
    idx = torch.arange(d2.shape[0] * K, dtype=torch.int32, device="cuda").view(-1, K)
    w = torch.from_numpy(z[f"w_{tag}"].reshape(-1)).cuda()
    raw = nat.logrho(d2, idx, K, 3, weights=w)
    got, ref = raw.cpu().numpy().astype(np.float64), z[f"raw_weighted_{tag}_d3"].astype(z[f"d2_{tag}"].dtype).astype(np.float64)
    assert np.isnan(got[ZERO_CORE]).all() and np.isnan(ref[ZERO_CORE]).all()
    fin = np.isfinite(ref)
    fin[KNIFE_EDGE] = False
    assert np.all(np.abs(got[fin] - ref[fin]) <= 10 * rtol * np.maximum(np.abs(ref[fin]), 1e-30))


def _cloud_with_copies(n, d, copies, second, dtype, seed=11):
    rng = np.random.default_rng(seed)
    P = rng.normal(0, 1, (n, d))
    P[:copies] = P[0]                           # `copies` coincident points
    P[copies:copies + second] = P[copies]       # a second, smaller stack
    return np.ascontiguousarray(P[rng.permutation(n)].astype(dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_search_and_fused_density_with_coincident_points(dtype):
    """>= k_den coincident points: their k-th distance is 0, the raw density NaN (reference semantics, :296-297), in the
    fused epilogue exactly as in the stand-alone pass; every other row is unaffected and still exact."""
    from astrolink_b200 import _native as nat
    from oracle import oracle
    nat.require_cuda()
    K = 12
    P = _cloud_with_copies(4000, 3, K + 5, K - 3, dtype)      # the second stack is smaller than k_den: positive core distance
    index = nat.Index(torch.from_numpy(P).cuda())
    d2, idx = index.knn(K)
    od2, oidx = oracle.knn(P, K)
    assert (nat.u32_numpy(idx) == oidx).all() and (d2.cpu().numpy() == od2).all(), "ties among coincident points break by index"
    raw_sep = nat.logrho(d2, idx, K, 3)
    raw_fused = torch.empty(P.shape[0], dtype=d2.dtype, device="cuda")
    link = torch.empty((P.shape[0], 5), dtype=torch.int32, device="cuda")
    index.knn(K, link_out=link, want_idx=False, logrho_out=raw_fused, d_intrinsic=3, want_d2=False)
    a, b = raw_sep.cpu().numpy(), raw_fused.cpu().numpy()
    zero_core = od2[:, K - 1] == 0
    assert zero_core.sum() == K + 5
    assert np.isnan(a[zero_core]).all() and np.isnan(b[zero_core]).all()
    assert np.isfinite(a[~zero_core]).all()
    assert (a.view(np.uint32 if dtype == np.float32 else np.uint64) == b.view(np.uint32 if dtype == np.float32 else np.uint64)).all()


@pytest.mark.gpu
def test_run_with_coincident_points_documented_behaviour():
    """Documented behaviour where the reference's is undefined (fastmath min/max over NaN, :305-313): NaN densities are
    ignored by the min/max and stay NaN; every stage completes; `ordering` is still a permutation and the group rows are
    well formed."""
    from astrolink_b200 import AstroLink
    P = _cloud_with_copies(6000, 3, 30, 27, np.float32)      # two stacks of >= k_den = 20 coincident points
    c = AstroLink(P)
    c.transform_data()
    c.estimate_density_and_kNN()
    lr = c.logRho.copy()
    assert int(np.isnan(lr).sum()) == 30 + 27
    fin = lr[np.isfinite(lr)]
    assert fin.min() == 0.0 and 0.999 < fin.max() <= 1.0
    c.aggregate()
    assert np.array_equal(np.sort(c.ordering), np.arange(P.shape[0], dtype=np.uint32))
    g = c.groups.astype(np.int64)
    assert (g[:, 0] < g[:, 1]).all() and (g[:, 1] < g[:, 2]).all() and (g[:, 2] <= P.shape[0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rescale_zero_variance_feature(dtype):
    """App. B Q6: the reference leaves the column of a zero-variance feature uninitialised (:239-242); here it is copied
    unchanged, the other features are divided by their population standard deviation as usual."""
    from astrolink_b200 import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(2)
    P = rng.normal(3, 2, (5000, 4)).astype(dtype)
    P[:, 2] = 7.25
    Pt, sd = nat.rescale(torch.from_numpy(P).cuda())
    Pt, sd = Pt.cpu().numpy(), sd.cpu().numpy()
    assert sd[2] == 0.0 and (Pt[:, 2] == P[:, 2]).all()
    for j in (0, 1, 3):
        ref = P[:, j].astype(np.float64) / P[:, j].astype(np.float64).std()
        assert np.allclose(Pt[:, j], ref, rtol=1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.gpu
def test_edges_when_rows_lack_their_own_point():
    """App. B Q7: with >= k_link coincident duplicates a neighbour row may not contain its own point; the reference then
    writes past the row's edge slots (:368-380).  Here every row still emits exactly k_link - 1 edges: the LAST entry is
    dropped.  Checked against the oracle, which documents the same rule."""
    from astrolink_b200 import _native as nat
    from oracle import oracle
    nat.require_cuda()
    rng = np.random.default_rng(4)
    n, L = 3000, 6
    # rows of distinct ids, none equal to the row's own point ...
    knn = (np.arange(n)[:, None] + 1 + np.stack([rng.permutation(n - 1)[:L] for _ in range(n)])) % n
    assert (knn != np.arange(n)[:, None]).all()
    knn = knn.astype(np.uint32)
    knn[::3, 0] = np.arange(0, n, 3)          # ... except every third row, which holds itself first, as a search result does
    lr = rng.random(n).astype(np.float32)
    w, pairs = nat.make_edges(torch.from_numpy(lr).cuda(), torch.from_numpy(knn.view(np.int32)).cuda(), L)
    opairs, oedges = oracle.make_graph(lr, knn)
    assert (nat.u32_numpy(pairs) == opairs).all() and (w.cpu().numpy() == oedges).all()
    assert pairs.shape[0] == n * (L - 1)
