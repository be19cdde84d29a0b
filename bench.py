#!/usr/bin/env python
"""Benchmark of the AstroLink hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

A step is one full `AstroLink(P).run()` over the synthetic 10M x 6-D phase-space
halo of BASELINE.json (configs[3], "C4"; it fits one GPU, so N=1 runs it whole
and N>1 shards its queries: strong scaling).  Metric: points/sec.

  value : points / (mean step time), P already resident in HBM when the step starts
  e2e   : the same through the public API from a pinned HOST array: H2D of P and
          D2H of logRho / ordering / groups / prominences inside the timed region
  roofline     : the kNN kernel (dominant kernel), FP32 pipe: 3*d flop per evaluated
                 (query, candidate) pair, pairs counted by the kernel, duration by CUDA
                 events on the launching stream; peak = FP32 FMA rate measured in-run
  cpu_baseline : the CPU oracle port of the same path timed on this box's host cores
                 on a bounded sample of the workload (rank 0, N=1 only)

Timing: W >= 3 warm-up steps, K timed steps bracketed by barrier + synchronize,
CUDA events, max over ranks.  The working set (P 240 MB, neighbour lists 1.6 GB,
60M edges) is far larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "points/sec for full run() (10M pts, 6D)"
UNIT = "points/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", help="C4 (default, 10M x 6D) or C2/C3 or n=<points> for a reduced halo")
    ap.add_argument("--cpu-sample", type=int, default=400000, help="points in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_workload(name):
    from astrolink_b200 import synth
    if name.startswith("n="):
        n = int(name[2:])
        return synth.halo6d(n, seed=4, n_sub=256, nested=True), {}, f"synthetic 6-D halo, n={n} (reduced C4)"
    w = synth.WORKLOADS[name]
    return w["make"](), dict(w["kwargs"]), w["desc"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm
def cpu_hot_path_seconds(P, kwargs):
    """One full run() of the CPU port: oracle hot path (all host threads) + the Python host statistics."""
    from oracle import oracle
    from astrolink_b200 import stats
    d = P.shape[1]
    k_den = 20
    k_link = max(int(np.ceil(11.97 * d ** (-2.23) - 22.97 * k_den ** (-0.57) + 10.03)), 7)
    t0 = time.perf_counter()
    r = oracle.hot_path(P, k_den=k_den, k_link=k_link, adaptive=kwargs.get("adaptive", 1))
    pFit, _ = stats.fit_prominence_model(r["prominences"][:, 1])
    sig = stats.significances_of(r["prominences"], pFit)
    S = stats.auto_threshold(r["prominences"].shape[0])
    stats.extract(r["groups"], sig, S, 1, P.shape[0])
    return time.perf_counter() - t0, r["times"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    P, kwargs, desc = make_workload(args.workload)
    n_s = min(args.cpu_sample * 2, P.shape[0])
    sample = np.ascontiguousarray(P[:n_s])          # the cloud is shuffled: a prefix is a uniform sample
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_hot_path_seconds(sample[: max(20000, n_s // 20)], kwargs)
    times = []
    stage = None
    for _ in range(args.steps):
        t, stage = cpu_hot_path_seconds(sample, kwargs)
        times.append(t)
    sec = float(np.mean(times))
    val = n_s / sec
    cores = oracle.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample": f"first {n_s} points of the shuffled cloud per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_s}-point prefix of the workload; oracle C port (OpenMP kNN, serial aggregate) + Python host stats; "
                                   f"stage seconds {json.dumps({k: round(v, 3) for k, v in stage.items()})}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    from astrolink_b200 import AstroLink, _native as nat
    from astrolink_b200 import dist as adist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    D = None
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
        D = adist.TorchDist()

    def barrier():
        if D is not None:
            D.barrier()
        torch.cuda.synchronize(dev)

    P, kwargs, desc = make_workload(args.workload)
    n, d = P.shape
    # pinned host copy (the user's array for the e2e leg) and a resident device copy (the `value` leg)
    P_pin_t = torch.empty((n, d), dtype=torch.from_numpy(P[:1]).dtype, pin_memory=True)
    P_pin_t.copy_(torch.from_numpy(P))
    P_pin = P_pin_t.numpy()
    dP = P_pin_t.to(dev)
    torch.cuda.synchronize(dev)

    def one_run(resident):
        c = AstroLink(P_pin, dist=D, **kwargs)
        if resident:
            c._use_device_input(dP)
        c.run()
        if rank == 0:
            # touch every result the user reads (D2H happens inside run(); these are host arrays already)
            _ = (c.logRho[0], c.ordering[0], c.groups.shape, c.prominences.shape, c.clusters.shape)
        return c

    def timed(resident, steps):
        """K steps bracketed by barrier + synchronize; total by CUDA events (max over ranks), plus per-step events."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        l0 = nat.launch_count()
        e0.record()
        marks[0].record()
        last = None
        for i in range(steps):
            last = one_run(resident)
            marks[i + 1].record()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        per_step = [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(steps)]
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if D is not None:
            import torch.distributed as td
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item()), last, (nat.launch_count() - l0) / steps, per_step

    for _ in range(args.warmup):
        one_run(True)
    one_run(False)          # the e2e path (H2D from the pinned host array) is warmed up too
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_value, c, launches, steps_value = timed(True, args.steps)
    ms_e2e, c2, _, steps_e2e = timed(False, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (kNN), from the last timed resident run
    stage = c.stage_ms()
    pairs = int(c._knn_pairs[0].item())
    if D is not None:
        import torch.distributed as td
        t = torch.tensor([float(pairs), stage["knn"]], dtype=torch.float64, device=dev)
        tmax = t.clone()
        td.all_reduce(t, op=td.ReduceOp.SUM)
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
        pairs_total, knn_ms = int(t[0].item()), float(tmax[1].item())
    else:
        pairs_total, knn_ms = pairs, stage["knn"]
    if rank != 0:
        if D is not None:
            import torch.distributed as td
            td.destroy_process_group()
        return 0
    fp32_peak = nat.fp32_peak_tflops()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flops = pairs * 3.0 * d                      # this rank's launch
    achieved = flops / (stage["knn"] * 1e-3) / 1e12
    K, L = c.k_den, c.k_link
    E = n * (L - 1)
    esz = 4
    hbm = {   # algorithmic bytes per launch (DESIGN.md "Kernels")
        "density": n * K * esz + n * esz + 3 * n * esz,
        "edges": n * L * 4 + n * L * esz + n * esz + E * (esz + 8),
        "sort": E * 4 + E * 68 + E * 16,
    }
    stages_out = {k: round(v, 3) for k, v in stage.items()}
    others = {k: {"ms": round(stage[k], 3), "GB/s": round(hbm[k] / (stage[k] * 1e-3) / 1e9, 1),
                  "frac_of_measured_hbm": round(hbm[k] / (stage[k] * 1e-3) / 1e9 / hbm_peak, 3)} for k in hbm if k in stage}

    h2d = P.nbytes
    d2h = int(c2.logRho.nbytes + c2.ordering.nbytes + c2.groups.nbytes + c2.prominences.nbytes)
    line = {
        "metric": METRIC, "value": n / (ms_value * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_value, "ms_each_step": steps_value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "n": n, "d": d, "k_den": K, "k_link": L,
                   "parallelism": f"query-sharded x{world}, aggregation on rank 0" if world > 1 else "single GPU",
                   "l2": "working set (P 240 MB, kNN 1.6 GB, 60M edges) >> 126 MB L2; no flush needed",
                   "groups": int(c.groups.shape[0]), "clusters": int(c.clusters.shape[0])},
        "e2e": {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "ms_each_step": steps_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(round(launches)),
        "clocks": clocks,
        "roofline": {"kernel": "knn_kernel<float,6,20>", "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak if fp32_peak else None, "traffic": None,
                     "pairs_evaluated": pairs, "flop_per_pair": 3 * d, "kernel_ms": stage["knn"],
                     "brute_force_equiv_tflops": (float(n) * n * 3 * d) / (knn_ms * 1e-3) / 1e12 / world,
                     "peak_source": "FP32 FMA issue rate measured in this run (al_fp32_peak_tflops); MEASURED_PEAKS.json has no FP32 figure",
                     "ceiling_note": "SUB+FMA mix caps the flop fraction at 0.75 of FMA peak"},
        "stage_ms": stages_out,
        "host_timers_ms": {k: round(1e3 * getattr(c, k), 2) for k in ("_transformTime", "_logRhoTime", "_aggregateTime", "_regrTime", "_rejTime", "_totalTime")},
        "hbm_passes": others,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        n_s = min(args.cpu_sample, n)
        sample = np.ascontiguousarray(P[:n_s])
        cpu_hot_path_seconds(sample[:20000], kwargs)
        sec, st = cpu_hot_path_seconds(sample, kwargs)
        line["cpu_baseline"] = {"value": n_s / sec, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                                "sample": f"{n_s}-point prefix of the shuffled workload, one run ({sec:.1f} s); oracle C port "
                                          f"(OpenMP KD-tree kNN, serial aggregate) + Python host stats"}
    print(json.dumps(line))
    if D is not None:
        import torch.distributed as td
        td.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
