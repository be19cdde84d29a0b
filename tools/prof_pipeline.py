"""One full AstroLink(P).run() bracketed by cudaProfilerStart/Stop, after a warm-up run, for ncu launch lists:
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file out.csv python tools/prof_pipeline.py C4
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from astrolink_b200 import synth, AstroLink

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
w = synth.WORKLOADS[name]
P = w["make"]()
dP = torch.from_numpy(P).cuda()
for it in range(2):
    c = AstroLink(P, **w["kwargs"])
    c._use_device_input(dP)
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    c.run()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.stop()
print({k: round(v, 3) for k, v in c.stage_ms().items()}, "groups", c.groups.shape[0], "clusters", c.clusters.shape[0])
