"""Host-side profile of one AstroLink step on the C4 workload (cProfile + wall clock)."""
import cProfile, pstats, sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import synth, AstroLink

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
P = synth.halo6d(n, seed=4, n_sub=256, nested=True)
Pp = torch.empty(P.shape, dtype=torch.float32, pin_memory=True); Pp.copy_(torch.from_numpy(P)); Pn = Pp.numpy()
dP = Pp.cuda()
def one():
    t0 = time.perf_counter()
    c = AstroLink(Pn); c._use_device_input(dP)
    t1 = time.perf_counter()
    c.run()
    t2 = time.perf_counter()
    _ = (c.logRho[0], c.ordering[0], c.groups.shape, c.prominences.shape, c.clusters.shape)
    t3 = time.perf_counter()
    return c, (t1 - t0, t2 - t1, t3 - t2)
for _ in range(3): c, _t = one()
torch.cuda.synchronize()
prev = c
for i in range(3):
    t0 = time.perf_counter()
    c, tt = one()
    t1 = time.perf_counter()
    prev = None
    t2 = time.perf_counter()
    prev = c
    print("ctor %.1f run %.1f touch %.1f | whole %.1f  drop-prev %.1f ms | gc %s" % (1e3*tt[0], 1e3*tt[1], 1e3*tt[2], 1e3*(t1-t0), 1e3*(t2-t1), gc.get_count()))
pr = cProfile.Profile(); pr.enable()
c2, _ = one()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
