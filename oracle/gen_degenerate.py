"""Golden vectors for degenerate inputs, from the reference's own functions (build container only).

    python -m oracle.gen_degenerate

Only the stages whose result the reference DEFINES on such inputs get a golden: the raw log-density
(`_compute_logRho_njit` / `_compute_weighted_logRho_njit`, astrolink.py:293-303, plain @njit: IEEE semantics, so a
zero core distance gives NaN and an all-equal row gives -inf).  `_normalise_njit` (:305-313) and `_rescale_njit`
(:236-243) are compiled with fastmath, where NaN/inf handling is undefined, and `_make_graph_njit` (:355-382) runs
past the end of a row that does not contain its own point; for those the tests pin this package's documented
behaviour instead (tests/test_gpu_degenerate.py).
Writes tests/golden/degenerate/density_rows.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402


def rows(dtype, K=20, seed=7):
    rng = np.random.default_rng(seed)
    base = np.sort(rng.uniform(0.01, 4.0, (64, K)), axis=1)
    base[:, 0] = 0.0
    r = [base]
    r.append(np.zeros((4, K)))                                   # >= k_den coincident points: core distance 0 -> NaN
    half = np.zeros((4, K)); half[:, K // 2:] = np.sort(rng.uniform(0.5, 2.0, (4, K - K // 2)), axis=1)
    r.append(half)                                               # some coincident copies, positive core distance
    r.append(np.full((4, K), 1.7))                               # all neighbours at the core distance: k - k = 0 -> -inf
    tiny = np.sort(rng.uniform(0, 1, (4, K)), axis=1) * 1e-38    # denormal distances
    tiny[:, 0] = 0.0
    r.append(tiny)
    huge = np.sort(rng.uniform(0.5, 1, (4, K)), axis=1) * (1e30 if dtype == np.float32 else 1e300)
    r.append(huge)                                               # core^(d/2) overflows in float32 / float64 for large d
    return np.ascontiguousarray(np.concatenate(r, axis=0).astype(dtype))


def main():
    assert ref_loader.available(), "needs /root/reference (build container only)"
    ref = ref_loader.load()
    out = {}
    rng = np.random.default_rng(3)
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        d2 = rows(dt)
        out[f"d2_{tag}"] = d2
        for d_int in (1, 3, 6):
            raw = ref.AstroLink._compute_logRho_njit(d2, d2.shape[1], d_int)
            out[f"raw_{tag}_d{d_int}"] = np.asarray(raw)          # float64, as the reference returns it
        w = rng.uniform(0.5, 2.0, d2.shape)
        out[f"w_{tag}"] = w
        out[f"raw_weighted_{tag}_d3"] = np.asarray(ref.AstroLink._compute_weighted_logRho_njit(d2, w, 3))
    path = os.path.join(ROOT, "tests", "golden", "degenerate", "density_rows.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        if k.startswith("raw"):
            print(k, "nan", int(np.isnan(v).sum()), "-inf", int(np.isneginf(v).sum()), "+inf", int(np.isposinf(v).sum()))
    print("wrote", path)


if __name__ == "__main__":
    main()
