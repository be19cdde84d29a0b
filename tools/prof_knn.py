"""Profiling driver: builds the index on a reduced C4 workload and runs the kNN
kernel (and optionally the whole pipeline) so ncu can capture it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
P = synth.halo6d(n, seed=4, n_sub=256, nested=True)
dP = torch.from_numpy(P).cuda()
Pt, _ = nat.rescale(dP)
index = nat.Index(Pt)
for _ in range(3):
    d2, idx, pairs = index.knn(20, count_pairs=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d2, idx, pairs = index.knn(20, count_pairs=True); e1.record(); torch.cuda.synchronize()
st = pairs.cpu().numpy().astype(float)
ms = e0.elapsed_time(e1)
print(f"checksum idx {int(idx.to(torch.int64).sum().item())} d2 {float(d2.double().sum().item()):.9e}")
print(f"n={n} knn {ms:.3f} ms  pairs/query {st[0]/n:.0f}  TFLOP/s {st[0]*18/ms/1e9:.2f}")
if st[1] > 0:      # diagnostic counters: only maintained by -DKNN_STATS=1 builds
    L = n / 32
    print(f"per leaf: tiles fetched {st[1]/L:.1f} (dense {st[2]/L:.1f})  descents {st[3]/L:.1f}  box-test iterations {st[7]/L:.1f}  "
          f"flush rounds {st[6]/L:.1f}   per query: pushes {st[4]/n:.1f} inserts {st[5]/n:.1f}")
