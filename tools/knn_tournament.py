"""A/B of kNN kernel builds in ONE process (development aid): builds the index once per library (the index format is
the same), times the search on the full C4 cloud and prints a checksum of the rows.
usage: python tools/knn_tournament.py <n> <tag> [<tag> ...]   ('' or 'base' = the default library)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, _native as nat

n = int(sys.argv[1])
tags = sys.argv[2:] or ["base"]
here = os.path.dirname(nat.lib_path())
P = synth.halo6d(n, seed=4, n_sub=256, nested=True)
dP = torch.from_numpy(P).cuda()
for tag in tags:
    nat._lib = None
    nat._LIB_PATH = os.path.join(here, "libastrolink_b200.so" if tag in ("", "base") else f"libastrolink_b200_{tag}.so")
    Pt, _ = nat.rescale(dP)
    index = nat.Index(Pt)
    for _ in range(2):
        d2, idx = index.knn(20)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d2, idx = index.knn(20); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    _, _, pairs = index.knn(20, count_pairs=True)
    pairs = int(pairs[0].item())
    print(f"{tag:6s} knn {best:8.3f} ms  tiles/query {pairs / 32 / n:6.1f}  checksum idx {int(idx.to(torch.int64).sum().item())} d2 {float(d2.double().sum().item()):.9e}", flush=True)
    del index, d2, idx
