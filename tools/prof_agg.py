"""Profiling driver for the rank-0 tail (edges, sort, aggregation) on a reduced or full C4 workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import synth, AstroLink

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = synth.halo6d(n, seed=4, n_sub=256, nested=True)
for it in range(2):
    c = AstroLink(P)
    c.transform_data(); c.estimate_density_and_kNN()
    torch.cuda.synchronize()
    c.aggregate()
    torch.cuda.synchronize()
print({k: round(v, 2) for k, v in c.stage_ms().items()})
import time
t = time.perf_counter(); c.compute_significances(); t1 = time.perf_counter(); c.extract_clusters(); t2 = time.perf_counter()
print(f"host: fit {1e3*(t1-t):.1f} ms extract {1e3*(t2-t1):.1f} ms groups {c.groups.shape[0]}")
