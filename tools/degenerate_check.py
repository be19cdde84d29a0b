"""Robustness: degenerate inputs must not hang or fault (results may be NaN like the reference's)."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import AstroLink
warnings.filterwarnings("ignore")
rng = np.random.default_rng(0)
cases = {
    "all identical (1000x3)": np.zeros((1000, 3), np.float32),
    "two distinct values": np.repeat(np.array([[0., 0.], [1., 1.]], np.float32), 400, axis=0),
    "many duplicates": np.repeat(rng.normal(0, 1, (50, 3)).astype(np.float32), 40, axis=0),
    "n=21, k_den=20": rng.normal(0, 1, (21, 2)),
    "n=40 1-D": rng.normal(0, 1, (40, 1)).astype(np.float32),
    "constant column": np.concatenate([rng.normal(0, 1, (3000, 2)), np.ones((3000, 1))], axis=1).astype(np.float32),
    "huge offsets": (rng.normal(0, 1, (5000, 3)) + 1e6).astype(np.float32),
    "inf-free extreme scales": (rng.normal(0, 1, (5000, 3)) * np.array([1e-20, 1.0, 1e20])).astype(np.float64),
}
for name, P in cases.items():
    try:
        c = AstroLink(np.ascontiguousarray(P))
        c.transform_data(); c.estimate_density_and_kNN(); c.aggregate()
        torch.cuda.synchronize()
        ok_perm = np.unique(c.ordering).size == P.shape[0]
        msg = f"ordering permutation={ok_perm} groups={c.groups.shape[0]} nan_logrho={int(np.isnan(c.logRho).sum())}"
        try:
            c.compute_significances(); c.extract_clusters(); msg += f" clusters={c.clusters.shape[0]}"
        except Exception as e:
            msg += f" host-stats raised {type(e).__name__}"
        print(f"OK    {name}: {msg}")
    except Exception as e:
        print(f"RAISE {name}: {type(e).__name__}: {str(e)[:150]}")
print("done")
