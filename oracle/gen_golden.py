"""Pins the oracle against the reference and writes tests/golden/*.npz.

Run in the build container only (needs /root/reference and numba):

    python -m oracle.gen_golden

For every case below the REFERENCE's own functions are executed, imported in
place from /root/reference/astrolink/astrolink.py (oracle/ref_loader.py):
transform_data (:218), estimate_density_and_kNN (:245; its pykdtree call is
answered by the oracle kNN, see ref_loader), _make_graph_njit (:355), the
canonical replacement of the unstable sort at :341
(np.argsort(-edges, kind='stable')), _aggregate_njit_float{32,64}_uint32
(:384/:495), compute_significances (:828) and extract_clusters (:909).  Their
outputs are the golden vectors.  The oracle's C restatement of each stage is
then run on the same inputs and compared (bit-exact for integer outputs,
tolerance for floats); the comparison report is written to
tests/golden/PINNING.json.  kNN has no reference arithmetic on this machine
(pykdtree absent): it is cross-checked against scipy.spatial.cKDTree instead
and recorded as "unpinned".
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import oracle, ref_loader  # noqa: E402
from astrolink_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def cases():
    rng = np.random.default_rng(1234)
    out = {}
    out["readme_c1_f64"] = dict(P=synth.readme_example(0), kwargs={})
    out["gauss1d_f32"] = dict(P=synth.two_gaussians(3000, 1, 11, np.float32), kwargs=dict(S=5, k_link=20))
    out["gauss2d_f64"] = dict(P=synth.two_gaussians(3000, 2, 12, np.float64), kwargs={})
    out["halo6d_f32"] = dict(P=synth.halo6d(6000, seed=2, n_sub=12), kwargs={})
    out["gmm3d_f32"] = dict(P=synth.gmm3d_nested(5000, seed=3), kwargs=dict(adaptive=0))
    # exact distance ties and equal densities: integer lattice, float32
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), indexing="ij"), -1).reshape(-1, 2)
    out["lattice_ties_f32"] = dict(P=np.ascontiguousarray(g.astype(np.float32)), kwargs=dict(adaptive=0, k_den=9, k_link=5))
    # duplicates (fewer than k_link coincident copies) + clusters
    base = rng.normal(0, 1, (700, 3))
    dup = np.concatenate((base, base[:150], base[:60], rng.normal(6, 0.3, (300, 3))), axis=0)
    out["duplicates_f32"] = dict(P=np.ascontiguousarray(dup.astype(np.float32)), kwargs=dict(k_den=12, k_link=8))
    # disconnected k_link graph: far-apart tight blobs, small k_link -> fix-up path (:442-461)
    blobs = np.concatenate([rng.normal(c, 0.05, (m, 2)) for c, m in ((0, 40), (50, 25), (100, 25), (150, 9), (200, 9))], axis=0)
    out["disconnected_f64"] = dict(P=np.ascontiguousarray(blobs), kwargs=dict(adaptive=0, k_den=8, k_link=3))
    out["disconnected_f32"] = dict(P=np.ascontiguousarray(blobs.astype(np.float32)), kwargs=dict(adaptive=0, k_den=8, k_link=3))
    out["klink2_f32"] = dict(P=synth.two_gaussians(400, 2, 13, np.float32), kwargs=dict(k_link=2))
    P = synth.two_gaussians(800, 3, 14, np.float64)
    out["weights_f64"] = dict(P=P, kwargs=dict(weights=rng.uniform(0.5, 2.0, P.shape[0]), d_intrinsic=2))
    out["weights_f32"] = dict(P=P.astype(np.float32), kwargs=dict(weights=rng.uniform(0.5, 2.0, P.shape[0])))
    for d in (1, 2, 3, 4, 5, 7, 8):
        out[f"dims_d{d}_f32"] = dict(P=synth.two_gaussians(600, d, 20 + d, np.float32), kwargs={})
    out["kden40_f64"] = dict(P=synth.two_gaussians(900, 4, 31, np.float64), kwargs=dict(k_den=40, k_link=11, h_style=0))
    return out


def run_reference(ref, P, kwargs):
    c = ref.AstroLink(P, **kwargs)
    g = {}
    c.transform_data()
    g["P_transform"] = np.ascontiguousarray(c.P_transform).copy()
    d2, idx = oracle.knn(g["P_transform"], c.k_den)
    g["knn_d2"], g["knn_idx"] = d2, idx
    c.estimate_density_and_kNN()            # pykdtree call answered by the oracle kNN
    g["logRho"] = c.logRho.copy()
    g["kNN"] = c.kNN.copy()
    assert (g["kNN"] == idx[:, :c.k_link]).all()
    if c.weights is None:
        raw = c._compute_logRho_njit(d2, c.k_den, c.d_intrinsic)
    else:
        raw = c._compute_weighted_logRho_njit(d2, c.weights[idx], c.d_intrinsic)
    g["logRho_raw"] = np.asarray(raw).astype(P.dtype)
    pairs, edges = c._make_graph_njit(c.logRho, c.kNN)
    g["pairs"], g["edges"] = pairs.copy(), edges.copy()
    order = np.argsort(-edges, kind="stable")     # canonical order
    g["order"] = order.astype(np.int64)
    ordered_pairs = pairs[order]
    fn = c._aggregate_njit_float64_uint32 if c.logRho.dtype == np.float64 else c._aggregate_njit_float32_uint32
    c.ordering, c.groups, c.prominences = fn(c.logRho, ordered_pairs)
    del c.kNN
    g["ordering"], g["groups"], g["prominences"] = c.ordering.copy(), c.groups.copy(), c.prominences.copy()
    g["k_link"] = np.int64(c.k_link)
    g["k_den"] = np.int64(c.k_den)
    g["d_intrinsic"] = np.int64(c.d_intrinsic)
    if c.groups.shape[0] >= 4:
        c.compute_significances()
        c.extract_clusters()
        g["pFit"], g["groups_sigs"] = np.asarray(c.pFit, dtype=np.float64), c.groups_sigs.copy()
        g["S"] = np.float64(c.S)
        g["clusters"], g["ids"], g["significances"] = c.clusters.copy(), np.asarray(c.ids), c.significances.copy()
    return g


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(a) & np.isfinite(b)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-30)))


def pin_oracle(name, P, kwargs, g, report):
    r = {}
    k_den, k_link, d_int = int(g["k_den"]), int(g["k_link"]), int(g["d_intrinsic"])
    adaptive = kwargs.get("adaptive", 1)
    # a1 rescale (numba fastmath std is not bit-reproducible: tolerance)
    Pt = oracle.rescale(P)[0] if adaptive else P
    r["rescale_relerr"] = relerr(Pt, g["P_transform"])
    # a3 kNN: tree search vs brute force (same contract), and vs scipy cKDTree (independent)
    d2b, idxb = oracle.knn(g["P_transform"], k_den, brute=True)
    r["knn_tree_eq_brute"] = bool((idxb == g["knn_idx"]).all() and (d2b == g["knn_d2"]).all())
    from scipy.spatial import cKDTree
    _, si = cKDTree(g["P_transform"].astype(np.float64)).query(g["P_transform"].astype(np.float64), k=k_den)
    r["knn_rows_identical_to_scipy_ckdtree"] = float((si.astype(np.uint32) == g["knn_idx"]).all(axis=1).mean())
    r["knn_neighbour_sets_identical_to_scipy"] = float(
        (np.sort(si, axis=1) == np.sort(g["knn_idx"].astype(np.int64), axis=1)).all(axis=1).mean())
    # a4/a5 density
    w = kwargs.get("weights")
    raw = oracle.logrho(g["knn_d2"], k_den, d_int, idx=g["knn_idx"], weights=w)
    r["logrho_raw_relerr"] = relerr(raw, g["logRho_raw"])
    lr = oracle.normalise(g["logRho_raw"])
    r["logrho_norm_abserr"] = float(np.nanmax(np.abs(lr.astype(np.float64) - g["logRho"].astype(np.float64))))
    # a8 graph, a9 order, a10 aggregate: injected with the reference's own logRho / kNN
    pairs, edges = oracle.make_graph(g["logRho"], g["kNN"])
    r["pairs_exact"] = bool((pairs == g["pairs"]).all())
    r["edges_exact"] = bool((edges == g["edges"]).all())
    order = oracle.sort_edges(g["edges"])
    r["order_exact"] = bool((order == g["order"]).all())
    ordering, groups, prom = oracle.aggregate(g["logRho"], g["pairs"][g["order"]])
    r["ordering_exact"] = bool((ordering == g["ordering"]).all())
    r["groups_exact"] = bool(groups.shape == g["groups"].shape and (groups == g["groups"]).all())
    r["prominences_exact"] = bool(prom.shape == g["prominences"].shape and (prom == g["prominences"]).all())
    r["prominences_maxabs"] = float(np.max(np.abs(prom.astype(np.float64) - g["prominences"].astype(np.float64)))) if prom.shape == g["prominences"].shape and prom.size else 0.0
    r["n"], r["d"], r["dtype"] = int(P.shape[0]), int(P.shape[1]), str(P.dtype)
    r["n_groups"] = int(g["groups"].shape[0])
    r["n_clusters"] = int(g["clusters"].shape[0]) if "clusters" in g else None
    report[name] = r
    ok = r["knn_tree_eq_brute"] and r["pairs_exact"] and r["edges_exact"] and r["order_exact"] and \
        r["ordering_exact"] and r["groups_exact"] and r["prominences_maxabs"] < 1e-6
    return ok


def main():
    assert ref_loader.available(), "needs /root/reference (build container only)"
    ref = ref_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    report = {}
    all_ok = True
    for name, case in cases().items():
        P, kwargs = case["P"], case["kwargs"]
        g = run_reference(ref, P, kwargs)
        ok = pin_oracle(name, P, kwargs, g, report)
        all_ok &= ok
        kw = {("kw_" + k): (np.asarray(v)) for k, v in kwargs.items()}
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), P=P, **kw, **g)
        print(("ok   " if ok else "FAIL ") + name, {k: v for k, v in report[name].items() if not isinstance(v, bool) or not v})
    with open(os.path.join(GOLDEN, "PINNING.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print("oracle pinned against the reference:", all_ok)
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
