"""C5 (100M x 3-D) warm run with aggregation phase timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from astrolink_b200 import synth, AstroLink
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10**8
P = synth.cosmo3d(n, seed=5)
Pp = torch.empty(P.shape, dtype=torch.float32, pin_memory=True); Pp.copy_(torch.from_numpy(P)); Pn = Pp.numpy(); del P
for it in range(2):
    c = None
    c = AstroLink(Pn, adaptive=0)
    torch.cuda.synchronize(); t = time.perf_counter(); c.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"run {it}: {dt*1e3:.1f} ms", {k: round(v, 1) for k, v in c.stage_ms().items()},
          {k: round(1e3*getattr(c, k), 1) for k in ("_transformTime", "_logRhoTime", "_aggregateTime", "_regrTime", "_rejTime")}, flush=True)
print("groups", c.groups.shape[0], "clusters", c.clusters.shape[0], "peak mem GB", round(torch.cuda.max_memory_allocated()/2**30, 1))
