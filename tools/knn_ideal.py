"""How many (query, tile) evaluations would a search with PERFECT bounds need?  (development aid)
For a sample of leaves: the number of leaf boxes within each query's final k-th distance, against the number the kernel
evaluates (pairs_evaluated / 32 / n).  usage: python tools/knn_ideal.py [n] [sample_leaves]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = 20
P = synth.halo6d(n, seed=4, n_sub=256, nested=True)
Pt, _ = nat.rescale(torch.from_numpy(P).cuda())
index = nat.Index(Pt)
d2, idx, pairs = index.knn(K, count_pairs=True)
pairs = int(pairs[0].item())
perm = index.perm().to(torch.int64)                      # position -> point (-1: padding)
npos = perm.shape[0]
valid = perm >= 0
pts = Pt[perm.clamp(min=0)]
d = Pt.shape[1]
big = torch.finfo(torch.float32).max
lo = torch.where(valid[:, None], pts, torch.full_like(pts, big)).view(-1, 32, d).amin(1)
hi = torch.where(valid[:, None], pts, torch.full_like(pts, -big)).view(-1, 32, d).amax(1)
nl = lo.shape[0]
g = torch.Generator(device="cpu").manual_seed(1)
leaves = torch.randperm(nl, generator=g)[:S].cuda()
lvl_ideal = [0, 0, 0, 0, 0]
tot_ideal = 0
tot_union = 0
nq = 0
for lf in leaves.tolist():
    pos = torch.arange(lf * 32, lf * 32 + 32, device="cuda")
    ok = valid[pos]
    q = pts[pos][ok]                                     # [m, d]
    kth = d2[perm[pos][ok], K - 1]                       # final k-th squared distance per query
    gap = torch.clamp(torch.maximum(lo[None] - q[:, None], q[:, None] - hi[None]), min=0)      # [m, nl, d]
    bd = (gap * gap).sum(-1)
    need = bd <= kth[:, None]
    tot_ideal += int(need.sum())
    ids = torch.arange(nl, device="cuda")
    for a in range(5):      # level of the lowest common ancestor with the own leaf (0 = the own leaf itself)
        same = (ids >> (5 * a)) == (lf >> (5 * a))
        below = (ids >> (5 * (a - 1))) == (lf >> (5 * (a - 1))) if a > 0 else torch.zeros_like(same)
        lvl_ideal[a] += int(need[:, same & ~below].sum())
    tot_union += int(need.any(0).sum())
    nq += int(ok.sum())
print(f"n {n} leaves {nl}; kernel: {pairs / 32 / n:.1f} (query, tile) evaluations per query")
print(f"perfect bounds: {tot_ideal / nq:.1f} tiles per query; tiles needed by at least one query of a leaf: {tot_union / S:.1f} (kernel fetches ~476 per leaf)")
print("perfect bounds, by level of the common ancestor (own leaf, parent, ...):", [round(v / nq, 1) for v in lvl_ideal])
if os.environ.get("AL_LIB_STATS"):
    nat._lib = None
    nat._LIB_PATH = os.environ["AL_LIB_STATS"]
    index2 = nat.Index(Pt)
    _, _, st = index2.knn(K, count_pairs=True)
    st = [int(v) for v in st.tolist()]
    print("kernel, needy queries of the evaluated tiles by ancestor level (parent, ...):", [round(v / n, 1) for v in st[1:5]])
    print("kernel, evaluated tiles per leaf needed by 1 / 2 / 3-4 queries:", [round(v / nl, 1) for v in st[5:8]])
