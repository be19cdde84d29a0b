"""Times al_sort_edges on a workload's edge list and prints a digest of the sorted pairs (development aid; AL_SORT_DIRECT=1 selects
the variant that carries the 64-bit pairs through the sort)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, AstroLink, _native as nat

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
w = synth.WORKLOADS[name]
P = w["make"]()
c = AstroLink(P, **w["kwargs"]); c._use_device_input(torch.from_numpy(P).cuda())
c.transform_data(); c.estimate_density_and_kNN()
wgt, pairs = nat.make_edges(c._dev("logRho"), c._dev("kNN"), c.k_link)
for _ in range(2): sp = nat.sort_edges(wgt, pairs)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sp = nat.sort_edges(wgt, pairs); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
dig = hashlib.sha1(sp.cpu().numpy().tobytes()).hexdigest()[:16]
print(f"{name} sort {best:.3f} ms  direct={os.environ.get('AL_SORT_DIRECT', '0')}  digest {dig}")
