"""Index build + kNN time for a given AL_TREE_LEVELS on C4 / C3-like data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, _native as nat
which = sys.argv[1] if len(sys.argv) > 1 else "halo"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000000
P = synth.halo6d(n, seed=4, n_sub=256, nested=True) if which == "halo" else synth.gmm3d_nested(n, seed=3) if which == "gmm" else synth.cosmo3d(n, seed=5)
dP = torch.from_numpy(P).cuda()
Pt = nat.rescale(dP)[0] if which == "halo" else dP
def t(f):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = f(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b), r
tbs, tks = [], []
index = d2 = idx = st = None
for _ in range(5):
    index = d2 = idx = st = None          # release first: a live previous result forces fresh cudaMallocs into the timing
    tb, index = t(lambda: nat.Index(Pt))
    tk, (d2, idx, st) = t(lambda: index.knn(20, count_pairs=True))
    tbs.append(tb); tks.append(tk)
tb, tk = min(tbs[1:]), min(tks[1:])
st = st.cpu().numpy().astype(float)
print(f"{which} n={n} levels={os.environ.get('AL_TREE_LEVELS','default')}: build {tb:.2f} ms  knn {tk:.2f} ms  pairs/query {st[0]/n:.0f} tiles/leaf {st[1]/(n/32):.1f}"
      f" eval/leaf {st[2]/(n/32):.1f} needy/tile {st[3]/max(st[1],1):.1f} pushes/q {st[4]/n:.1f} inserts/q {st[5]/n:.1f} merge_iter/leaf {st[6]/(n/32):.1f}")
