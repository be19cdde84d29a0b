"""Multi-GPU path on real GPUs: the query-sharded run must reproduce the
single-GPU results bit for bit (skipped with fewer than 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["AL_ROOT"])
import numpy as np, torch, torch.distributed as td
from astrolink_b200 import AstroLink, synth
from astrolink_b200 import dist as adist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
P = synth.halo6d(150001, seed=9, n_sub=32)
s = None
if rank == 0:
    s = AstroLink(P)
    s.run()
# peer=True: kernels store into rank 0's buffers over NVLink (run twice: cached mappings); peer=False: NCCL all-gather
for D in (adist.TorchDist(peer=True), adist.TorchDist(peer=False)):
    for rep in range(2):
        c = AstroLink(P, dist=D)
        c.run()
        if rank == 0:
            assert (c.logRho == s.logRho).all(), "logRho differs between sharded and single-GPU runs"
            assert (c.ordering == s.ordering).all() and (c.groups == s.groups).all()
            assert (c.clusters == s.clusters).all() and list(c.ids) == list(s.ids)
# weighted densities (a5) with n % 32 != 0: padding rows of the last leaf must not be gathered through
w = np.random.default_rng(3).uniform(0.5, 2.0, P.shape[0])
sw = None
if rank == 0:
    sw = AstroLink(P, weights=w)
    sw.run()
for D in (adist.TorchDist(peer=True), adist.TorchDist(peer=False)):
    c = AstroLink(P, weights=w, dist=D)
    c.run()
    if rank == 0:
        assert (c.logRho == sw.logRho).all() and (c.ordering == sw.ordering).all() and (c.groups == sw.groups).all()
if rank == 0:
    print("MULTI_OK", c.clusters.shape[0])
td.destroy_process_group()
'''

# The same sharded code on ONE GPU: `world` processes share cuda:0 (gloo process group for the barrier and the handle
# broadcast; the peer buffers are CUDA IPC mappings of rank 0's memory, here on the same device).  world 2: rank 0 takes
# 3 of 8 chunks; world 3: 2 of 12, or -- AL_TAIL_WIDTHS=0,1,1 -- none: a dedicated tail rank that gets the index permutation from rank 1.
SCRIPT_ONE_GPU = r'''
import os, sys
sys.path.insert(0, os.environ["AL_ROOT"])
import numpy as np, torch, torch.distributed as td
from astrolink_b200 import AstroLink, synth
from astrolink_b200 import dist as adist
rank = int(os.environ["RANK"])
torch.cuda.set_device(0)
td.init_process_group("gloo")
for P, kw in ((synth.halo6d(150001, seed=9, n_sub=32), {}), (synth.gmm3d_nested(60013, seed=4), dict(adaptive=0)),
              (synth.two_gaussians(20001, 2, 3, np.float64), {})):
    s = None
    w = np.random.default_rng(3).uniform(0.5, 2.0, P.shape[0])
    D = adist.TorchDist(peer=True)
    for weights in (None, w):
        if rank == 0:
            s = AstroLink(P, weights=weights, **kw)
            s.run()
        for rep in range(3):      # three runs: both halves of the double-buffered peer arrays, cached mappings
            c = AstroLink(P, weights=weights, dist=D, **kw)
            c.run()
            if rank == 0:
                assert (c.logRho == s.logRho).all(), "logRho differs between sharded and single-GPU runs"
                assert (c.ordering == s.ordering).all() and (c.groups == s.groups).all()
                assert (c.prominences == s.prominences).all()
                assert (c.clusters == s.clusters).all() and list(c.ids) == list(s.ids)
if rank == 0:
    print("MULTI_OK widths", D.widths)
td.barrier()
td.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_run_equals_single_gpu(tmp_path):
    script = tmp_path / "multi.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, AL_ROOT=ROOT)
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_OK" in r.stdout


@pytest.mark.parametrize("world,widths", [(2, None), (3, None), (3, "0,1,1")])
def test_sharded_run_on_one_gpu_equals_single_process(tmp_path, world, widths):
    """The sharded path without a second GPU: `world` processes on cuda:0 (see SCRIPT_ONE_GPU).  widths None: the default
    tail-aware deal (rank 0 searches a smaller share); "0,1,1": rank 0 as a dedicated tail rank (the deal from five GPUs
    up), which builds no index and receives the index permutation from rank 1."""
    script = tmp_path / "multi1.py"
    script.write_text(SCRIPT_ONE_GPU)
    env = dict(os.environ, AL_ROOT=ROOT, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
    if widths is not None:
        env["AL_TAIL_WIDTHS"] = widths
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29741 + world + (10 if widths else 0)), str(script)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_OK" in r.stdout
