"""Dense inner loop of the kNN kernel in isolation: a KNN_BRUTE build (no pruning, every tile evaluated by
every warp with the packed f32x2 path) on a small uniform cloud.  Reports FP32 TFLOP/s (3*d flop per pair)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import _native as nat
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, int(sys.argv[2]) if len(sys.argv) > 2 else 6
P = np.random.default_rng(0).uniform(0, 1, (n, d)).astype(np.float32)
index = nat.Index(torch.from_numpy(P).cuda())
for _ in range(2): index.knn(20)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); d2, idx, st = index.knn(20, count_pairs=True); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b); pairs = float(st[0].item())
peak = nat.fp32_peak_tflops()
print(f"dense mode n={n} d={d}: {ms:.3f} ms, {pairs:.3e} pairs ({pairs/n/n:.2f} n^2), {pairs*3*d/ms/1e9:.2f} TFLOP/s = {pairs*3*d/ms/1e9/peak*100:.1f}% of measured FP32 FMA peak {peak:.1f} (SUB+FMA ceiling 75%)")
