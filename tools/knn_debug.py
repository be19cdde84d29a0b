"""Development aid: kNN vs the CPU oracle on one random cloud, with details of the first mismatching rows.
usage: python tools/knn_debug.py n d k [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import _native as nat
from oracle import oracle

n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rep = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(n + d)
P = np.concatenate([rng.normal(0, 1, (n // 2, d)), rng.normal(3, 0.2, (n - n // 2, d))]).astype(np.float32)
P = np.ascontiguousarray(P[rng.permutation(n)])
od2, oidx = oracle.knn(P, k)
index = nat.Index(torch.from_numpy(P).cuda())
perm = index.perm().cpu().numpy()
inv = np.empty(n, dtype=np.int64); inv[perm[perm >= 0]] = np.nonzero(perm >= 0)[0]
for r in range(rep):
    d2, idx = index.knn(k)
    torch.cuda.synchronize()
    gi, gd = nat.u32_numpy(idx), d2.cpu().numpy()
    bad = np.nonzero((gi != oidx).any(axis=1) | (gd != od2).any(axis=1))[0]
    print(f"run {r}: lib={os.path.basename(nat.lib_path())} n={n} d={d} k={k}: {bad.size} bad rows; leaves hit: {np.unique(inv[bad] // 32).size}")
    for i in bad[:3]:
        cols = np.nonzero((gi[i] != oidx[i]) | (gd[i] != od2[i]))[0]
        print(f"  row {i} (position {inv[i]}, lane {inv[i] % 32}): cols {cols.tolist()}")
        print(f"     got ids {gi[i][cols].tolist()} d2 {gd[i][cols].tolist()}")
        print(f"     exp ids {oidx[i][cols].tolist()} d2 {od2[i][cols].tolist()}")
