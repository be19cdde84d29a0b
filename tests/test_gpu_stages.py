"""GPU parity of each stage of the hot path, called through the C ABI, against
the committed golden vectors (reference outputs) and the CPU oracle."""
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


@pytest.mark.parametrize("name", golden_names())
def test_knn_bit_exact_vs_golden(name, golden):
    nat = _native()
    g = golden(name)
    Pt = _cu(g["P_transform"])
    k = int(g["k_den"])
    index = nat.Index(Pt)
    d2, idx = index.knn(k)
    torch.cuda.synchronize()
    assert (nat.u32_numpy(idx) == g["knn_idx"]).all(), "kNN indices must be bit-exact (ties by index)"
    assert (d2.cpu().numpy() == g["knn_d2"]).all(), "squared distances must be bit-exact (same fma chain)"


@pytest.mark.parametrize("name", golden_names())
def test_knn_tree_order_rows(name, golden):
    nat = _native()
    g = golden(name)
    Pt = _cu(g["P_transform"])
    k = int(g["k_den"])
    index = nat.Index(Pt)
    d2, idx = index.knn(k, row_order=1)
    perm = index.perm().cpu().numpy()
    valid = perm >= 0
    assert sorted(perm[valid].tolist()) == list(range(Pt.shape[0]))
    got = nat.u32_numpy(idx)[valid]
    assert (got == g["knn_idx"][perm[valid]]).all()


@pytest.mark.parametrize("name", golden_names())
def test_density_vs_golden(name, golden):
    nat = _native()
    g = golden(name)
    w = _cu(g["kw_weights"].astype(np.float64)) if "kw_weights" in g else None
    d2, idx = _cu(g["knn_d2"]), _cu(g["knn_idx"].view(np.int32))
    raw = nat.logrho(d2, idx, int(g["k_den"]), int(g["d_intrinsic"]), weights=w)
    ref = g["logRho_raw"].astype(np.float64)
    got = raw.cpu().numpy().astype(np.float64)
    fin = np.isfinite(ref)
    assert (np.isfinite(got) == fin).all()
    # north_star tolerance: 1e-5 relative
    assert np.all(np.abs(got[fin] - ref[fin]) <= 1e-5 * np.maximum(np.abs(ref[fin]), 1e-30))
    lr = nat.normalise_(raw.clone())
    refn = g["logRho"].astype(np.float64)
    gotn = lr.cpu().numpy().astype(np.float64)
    fin = np.isfinite(refn)
    assert np.all(np.abs(gotn[fin] - refn[fin]) <= 1e-5 * np.maximum(np.abs(refn[fin]), 1.0) )


@pytest.mark.parametrize("name", golden_names())
def test_edges_and_order_bit_exact(name, golden):
    nat = _native()
    g = golden(name)
    L = int(g["k_link"])
    lr = _cu(g["logRho"])
    kNN = _cu(g["kNN"].view(np.int32))
    w, pairs = nat.make_edges(lr, kNN, L)
    assert (w.cpu().numpy() == g["edges"]).all()
    assert (nat.u32_numpy(pairs) == g["pairs"]).all()
    sp, order = nat.sort_edges(w, pairs, want_order=True)
    assert (nat.u32_numpy(order).astype(np.int64) == g["order"]).all(), "canonical order: weight desc, slot asc"
    assert (nat.u32_numpy(sp) == g["pairs"][g["order"]]).all()
    # without the slot order the pairs ride through the sort as 64-bit values: the same stable order, the same bytes
    assert (nat.u32_numpy(nat.sort_edges(w, pairs)) == g["pairs"][g["order"]]).all()


@pytest.mark.parametrize("name", [n for n in golden_names()])
def test_rescale_vs_golden(name, golden):
    nat = _native()
    g = golden(name)
    if "kw_adaptive" in g and int(g["kw_adaptive"]) == 0:
        pytest.skip("adaptive=0")
    P = _cu(g["P"])
    Pt, std = nat.rescale(P)
    ref = g["P_transform"].astype(np.float64)
    got = Pt.cpu().numpy().astype(np.float64)
    tol = 1e-6 if g["P"].dtype == np.float32 else 1e-13
    assert np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), 1e-30))


@pytest.mark.parametrize("n,d,dtype,k", [(20000, 3, np.float32, 20), (30011, 6, np.float32, 20), (5000, 2, np.float64, 9),
                                         (12345, 6, np.float64, 20), (40000, 1, np.float32, 18), (9000, 8, np.float32, 33),
                                         (3000, 4, np.float32, 100), (33, 3, np.float32, 5), (31, 2, np.float32, 31), (6000, 11, np.float32, 20), (4000, 16, np.float64, 12),
                                         (5000, 9, np.float64, 20), (7000, 16, np.float32, 25),
                                         (100000, 6, np.float32, 20)])
def test_knn_vs_oracle_random(n, d, dtype, k):
    from oracle import oracle
    nat = _native()
    rng = np.random.default_rng(n + d)
    P = np.concatenate([rng.normal(0, 1, (n // 2, d)), rng.normal(3, 0.2, (n - n // 2, d))]).astype(dtype)
    P = np.ascontiguousarray(P[rng.permutation(n)])
    index = nat.Index(_cu(P))
    d2, idx, pairs = index.knn(k, count_pairs=True)
    od2, oidx = oracle.knn(P, k)
    assert (nat.u32_numpy(idx) == oidx).all()
    assert (d2.cpu().numpy() == od2).all()
    assert int(pairs[0].item()) >= n * min(n, 32)


@pytest.mark.parametrize("a,b", [(1.3, 12.0), (1.0, 1.0), (2.7, 40.0), (1.05, 3.0), (8.0, 2.5)])
def test_significance_kernel_vs_scipy(a, b):
    """al_significance == scipy's norm.isf(beta.sf(x, a, b)) (astrolink.py:858), including the far tails."""
    from scipy.stats import beta, norm
    nat = _native()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.beta(a, b, 20000), rng.uniform(0, 1, 20000), [0.0, 1e-300, 1e-12, 0.5, 1 - 1e-12, 0.999999, 1.0],
                        np.linspace(0.9, 0.9999, 200)])
    ref = norm.isf(beta.sf(x, a, b))
    got = nat.significance(_cu(x), a, b).cpu().numpy()
    fin = np.isfinite(ref)
    assert (np.isfinite(got) == fin).all() and (got[~fin] == ref[~fin]).all()
    assert np.allclose(got[fin], ref[fin], rtol=1e-9, atol=1e-9)
    x32 = x.astype(np.float32)
    ref32 = norm.isf(beta.sf(x32.astype(np.float64), a, b))
    got32 = nat.significance(_cu(x32), a, b).cpu().numpy()
    fin = np.isfinite(ref32)
    assert np.allclose(got32[fin], ref32[fin], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("G,dtype", [(50000, np.float32), (20001, np.float64), (300000, np.float32)])
def test_w1_objective_kernel_vs_numpy(G, dtype):
    """al_w1_prepare / al_w1_eval == the NumPy objective of astrolink_b200.stats (astrolink.py:896-907),
    and the fit driven by it lands on the same parameters."""
    from astrolink_b200 import stats
    nat = _native()
    rng = np.random.default_rng(G)
    prom = np.stack([rng.beta(2.0, 9.0, G), rng.beta(1.4, 11.0, G)], axis=1).astype(dtype)
    prom[:5, 1] = 0.0                      # zero-width gaps
    obj = nat.W1Objective(_cu(prom))
    ps = np.sort(prom[:, 1].astype(np.float64))
    assert np.array_equal(obj.sorted, ps)
    ref = stats._W1Objective(ps)
    for p in ([1.4, 11.0], [1.0, 1.0], [2.5, 30.0], [1.05, 4.0]):
        a, b = ref(np.array(p)), obj(np.array(p))
        assert abs(a - b) <= 1e-10 * max(1.0, abs(a)), (p, a, b)
    f_ref, ok1 = stats.fit_prominence_model(prom[:, 1])
    f_gpu, ok2 = stats.fit_prominence_model(None, objective=obj, p_sorted=obj.sorted)
    assert ok1 and ok2 and np.allclose(f_ref, f_gpu, rtol=1e-5)


@pytest.mark.parametrize("n,d,world,chunk_leaves", [(30011, 6, 3, 8), (20000, 3, 2, 64), (9000, 2, 4, 4), (30011, 6, 8, 1024)])
def test_knn_dealt_launches_equal_one_launch(n, d, world, chunk_leaves):
    """The sharded stage on one GPU: the `world` dealt launches (al_knn_dealt + al_logrho_dealt, rows sent in
    index-position order, un-permuted with al_scatter_rows) reproduce the single launch bit for bit, and the
    k_link columns written by the epilogue (al_knn_link) equal the first columns of the full rows."""
    from astrolink_b200 import dist as adist
    nat = _native()
    rng = np.random.default_rng(n + d)
    P = rng.normal(size=(n, d)).astype(np.float32)
    P[: n // 3] *= 0.05                                  # a dense clump: regional differences in search cost
    index = nat.Index(_cu(P))
    K, L = 20, 7
    link = torch.full((n, L), -7, dtype=torch.int32, device="cuda")
    d2_ref, idx_ref = index.knn(K, link_out=link)
    assert (link == idx_ref[:, :L]).all()
    raw_ref = nat.logrho(d2_ref, idx_ref, K, d)
    # link columns only (no full index rows)
    link2 = torch.full((n, L), -7, dtype=torch.int32, device="cuda")
    d2_b, idx_b = index.knn(K, link_out=link2, want_idx=False)
    assert idx_b is None and (link2 == link).all() and (d2_b == d2_ref).all()

    positions = index.positions
    knn_pos = torch.full((positions, L), -5, dtype=torch.int32, device="cuda")
    raw_pos = torch.full((positions,), float("nan"), dtype=torch.float32, device="cuda")
    covered = 0
    for r in range(world):
        begin, chunk, stride, rows = adist.deal_chunks(positions, r, world, nat.LEAF, chunk_leaves)
        d2, idx = index.knn(K, pos_begin=begin, pos_end=positions, row_order=1, link_out=knn_pos, want_idx=False,
                            chunk=chunk, stride=stride)
        assert d2.shape[0] == rows and idx is None
        nat.logrho_dealt(d2, None, K, d, begin, positions, chunk, stride, raw_pos)
        covered += rows
    assert covered >= positions
    perm = index.perm()
    kNN = torch.full((n, L), -9, dtype=torch.int32, device="cuda")
    raw = torch.full((n,), float("nan"), dtype=torch.float32, device="cuda")
    nat.scatter_rows(knn_pos, perm, kNN)
    nat.scatter_rows(raw_pos.view(-1, 1), perm, raw.view(-1, 1))
    torch.cuda.synchronize()
    assert (kNN == idx_ref[:, :L]).all(), "dealt launches must reproduce the single launch"
    assert (raw == raw_ref).all()


@pytest.mark.parametrize("n,d,dtype,k,d_int", [(30011, 6, np.float32, 20, 6), (20000, 3, np.float32, 20, 2), (9000, 2, np.float64, 9, 2),
                                               (5000, 8, np.float64, 33, 8), (33, 3, np.float32, 5, 3), (7000, 16, np.float32, 25, 5)])
def test_fused_density_equals_separate_pass(n, d, dtype, k, d_int):
    """al_knn_density (a3 + a4 in one launch, astrolink.py:279 and :293-297): the raw log-density written by the search
    epilogue equals al_logrho of the same distance rows bit for bit -- by original index, and by index position for
    dealt launches -- and with d2_out == NULL nothing else changes."""
    from astrolink_b200 import dist as adist
    nat = _native()
    rng = np.random.default_rng(7 * n + d)
    P = np.concatenate([rng.normal(0, 1, (n // 2, d)), rng.normal(2, 0.1, (n - n // 2, d))]).astype(dtype)
    index = nat.Index(_cu(P))
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    L = min(7, k)
    d2_ref, idx_ref = index.knn(k)
    raw_ref = nat.logrho(d2_ref, idx_ref, k, d_int)
    # fused, by original index, distance rows still written
    link = torch.full((n, L), -7, dtype=torch.int32, device="cuda")
    raw = torch.full((n,), float("nan"), dtype=tdt, device="cuda")
    d2, idx = index.knn(k, link_out=link, logrho_out=raw, d_intrinsic=d_int)
    assert (d2 == d2_ref).all() and (idx == idx_ref).all() and (link == idx_ref[:, :L]).all()
    assert (raw.view(torch.int32 if dtype == np.float32 else torch.int64) ==
            raw_ref.view(torch.int32 if dtype == np.float32 else torch.int64)).all(), "fused density must equal al_logrho bit for bit"
    # fused, no distance rows, no index rows
    link2 = torch.full((n, L), -7, dtype=torch.int32, device="cuda")
    raw2 = torch.full((n,), float("nan"), dtype=tdt, device="cuda")
    d2n, idxn = index.knn(k, link_out=link2, want_idx=False, logrho_out=raw2, d_intrinsic=d_int, want_d2=False)
    assert d2n is None and idxn is None and (link2 == link).all() and (raw2 == raw_ref).all()
    # dealt launches: densities by index position
    positions = index.positions
    knn_pos = torch.full((positions, L), -5, dtype=torch.int32, device="cuda")
    raw_pos = torch.full((positions,), float("nan"), dtype=tdt, device="cuda")
    for r in range(3):
        begin, chunk, stride, rows = adist.deal_chunks(positions, r, 3, nat.LEAF, 8)
        index.knn(k, pos_begin=begin, pos_end=positions, row_order=1, link_out=knn_pos, want_idx=False, chunk=chunk, stride=stride,
                  logrho_out=raw_pos, d_intrinsic=d_int, want_d2=False)
    raw3 = torch.full((n,), float("nan"), dtype=tdt, device="cuda")
    kNN3 = torch.full((n, L), -9, dtype=torch.int32, device="cuda")
    nat.scatter_rows(raw_pos.view(-1, 1), index.perm(), raw3.view(-1, 1))
    nat.scatter_rows(knn_pos, index.perm(), kNN3)
    torch.cuda.synchronize()
    assert (raw3 == raw_ref).all() and (kNN3 == idx_ref[:, :L]).all()


def test_weighted_density_skips_padding_rows():
    """Position-ordered launches (multi-GPU path) fill the padding rows of the last leaf with 0xffffffff ids; the weighted
    density (a5, astrolink.py:299-303) must not gather weights through them (n % 32 != 0)."""
    nat = _native()
    n, d, k = 1003, 3, 12
    rng = np.random.default_rng(5)
    P = rng.normal(size=(n, d)).astype(np.float32)
    w = torch.from_numpy(rng.uniform(0.5, 2.0, n)).cuda()
    index = nat.Index(_cu(P))
    d2, idx = index.knn(k, row_order=1)
    assert d2.shape[0] == index.positions and index.positions > n
    raw = nat.logrho(d2, idx, k, d, weights=w)
    d2o, idxo = index.knn(k)
    raw_o = nat.logrho(d2o, idxo, k, d, weights=w)
    perm = index.perm().cpu().numpy()
    torch.cuda.synchronize()
    got = raw.cpu().numpy()[perm >= 0]
    assert (got == raw_o.cpu().numpy()[perm[perm >= 0]]).all()


@pytest.mark.parametrize("L,stride,dtype", [(7, 7, np.float32), (7, 20, np.float32), (9, 9, np.float64), (2, 2, np.float32),
                                            (18, 18, np.float32), (40, 40, np.float64)])
def test_make_edges_layouts_vs_oracle(L, stride, dtype):
    """a8 with contiguous and strided neighbour rows, short and long rows (staged kernel and its fallback), both dtypes."""
    from oracle import oracle
    nat = _native()
    n = 5003
    rng = np.random.default_rng(L * 100 + stride)
    lr = rng.random(n).astype(dtype)
    lr[rng.integers(0, n, 300)] = lr[7]                  # ties
    knn = np.empty((n, L), dtype=np.uint32)
    for i in range(n):
        others = rng.choice(n - 1, L - 1, replace=False)
        others[others >= i] += 1
        row = np.concatenate(([i], others))
        if i % 5 == 0:
            rng.shuffle(row)                             # the point itself is not always in column 0
        knn[i] = row
    pairs_o, edges_o = oracle.make_graph(lr, knn)
    full = np.zeros((n, stride), dtype=np.uint32)
    full[:, :L] = knn
    dk = torch.from_numpy(full.view(np.int32)).cuda()[:, :L]
    w, pairs = nat.make_edges(_cu(lr), dk, L)
    torch.cuda.synchronize()
    assert (w.cpu().numpy() == edges_o).all()
    assert (nat.u32_numpy(pairs) == pairs_o).all()
