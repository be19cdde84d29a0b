import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
log = open(f"gpurun_out/multi_rank{rank}.log", "w")
def say(*a):
    print(*a, file=log, flush=True)
try:
    from astrolink_b200 import AstroLink, synth
    from astrolink_b200 import dist as adist
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    say("init ok")
    D = adist.TorchDist(peer=True)
    t = D.peer.get("probe", (1000, 7), torch.int32, torch.device("cuda", local))
    say("peer buffer", t.shape, hex(t.data_ptr()), "owner", t.owner)
    from astrolink_b200 import _native as nat
    src = torch.full((10, 7), rank + 1, dtype=torch.int32, device=f"cuda:{local}")
    rows = torch.arange(rank * 10, rank * 10 + 10, dtype=torch.int32, device=f"cuda:{local}")
    nat.scatter_rows(src, rows, t)                 # a kernel on this rank's GPU storing into rank 0's memory
    torch.cuda.synchronize(); td.barrier()
    if rank == 0:
        say("probe rows", t.tensor[:20, 0].tolist())
    P = synth.halo6d(150001, seed=9, n_sub=32)
    c = AstroLink(P, dist=D)
    c.run()
    say("run ok", c.stage_ms())
    td.destroy_process_group()
except BaseException:
    say("FAILED\n" + traceback.format_exc()[-3000:])
    log.close()
    os._exit(1)
