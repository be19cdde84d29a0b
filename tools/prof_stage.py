"""Runs the hot path on a workload with cudaProfilerStart/Stop around ONE stage, so that
    ncu --profile-from-start off [--set full | --metrics ...] python tools/prof_stage.py <stage> [workload]
captures only that stage's kernels (after a complete warm-up run).
stages: rescale, index, knn, density (the stand-alone al_logrho pass on [n, k_den] rows), edges, sort, aggregate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, AstroLink, _native as nat

stage = sys.argv[1]
name = sys.argv[2] if len(sys.argv) > 2 else "C4"
w = synth.WORKLOADS[name]
P = w["make"]()
dP = torch.from_numpy(P).cuda()
warm = AstroLink(P, **w["kwargs"]); warm._use_device_input(dP); warm.run(); del warm
torch.cuda.synchronize()
prof = torch.cuda.profiler


def bracket(fn):
    torch.cuda.synchronize(); prof.start(); r = fn(); torch.cuda.synchronize(); prof.stop(); return r


c = AstroLink(P, **w["kwargs"]); c._use_device_input(dP)
if stage == "rescale":
    bracket(c.transform_data)
    sys.exit(0)
c.transform_data()
Pt = c._dev("P_transform")
K, L = c.k_den, c.k_link
if stage == "index":
    bracket(lambda: nat.Index(Pt))
elif stage == "knn":
    index = nat.Index(Pt)
    kNN = torch.empty((P.shape[0], L), dtype=torch.int32, device="cuda"); raw = torch.empty(P.shape[0], dtype=Pt.dtype, device="cuda")
    bracket(lambda: index.knn(K, link_out=kNN, want_idx=False, logrho_out=raw, d_intrinsic=c.d_intrinsic, want_d2=False))
elif stage == "density":
    index = nat.Index(Pt)
    d2, idx = index.knn(K)
    bracket(lambda: nat.logrho(d2, None, K, c.d_intrinsic))
else:
    c.estimate_density_and_kNN()
    lr, kNN = c._dev("logRho"), c._dev("kNN")
    if stage == "edges":
        bracket(lambda: nat.make_edges(lr, kNN, L))
    else:
        wgt, pairs = nat.make_edges(lr, kNN, L)
        if stage == "sort":
            bracket(lambda: nat.sort_edges(wgt, pairs))
        elif stage == "aggregate":
            sp = nat.sort_edges(wgt, pairs)
            del wgt, pairs
            bracket(lambda: nat.aggregate(lr, sp))
        else:
            raise SystemExit("unknown stage " + stage)
print("profiled stage", stage, "on", name)
