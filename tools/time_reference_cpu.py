"""Times the REFERENCE's own CPU path on this machine's cores (build container only: needs /root/reference).

    python tools/time_reference_cpu.py C3 C4 > profiles/r02_reference_numba_cpu.json

The reference is imported in place (never copied).  Its numba kernels and host stages run unmodified
(transform_data :218, estimate_density_and_kNN :245, aggregate :315 with its own unstable argsort at :341,
compute_significances :828, extract_clusters :909).  The one substitution: `pykdtree` (astrolink.py:16) is not
installed and not installable offline, so `KDTree(P).query(P, k, sqr_dists=True)` (:270, :279) is answered by
`scipy.spatial.cKDTree(P).query(P, k, workers=-1)` -- the stand-in SURVEY.md section 8(d) names.  JIT compilation is
excluded by a warm-up run on 20 000 points.  OMP/NUMBA thread counts are exported before the interpreter loads numba.
"""
import json
import os
import sys
import time
import types

NCPU = os.cpu_count() or 1
os.environ.setdefault("OMP_NUM_THREADS", str(NCPU))
os.environ.setdefault("NUMBA_NUM_THREADS", str(NCPU))

import importlib.util  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_FILE = "/root/reference/astrolink/astrolink.py"


class _CKDTreeStub:
    def __init__(self, data, leafsize=16):
        from scipy.spatial import cKDTree
        self.data = np.ascontiguousarray(data)
        self.tree = cKDTree(self.data, leafsize=leafsize)

    def query(self, q, k=1, sqr_dists=False, **kw):
        d, i = self.tree.query(q, k=k, workers=-1)
        d = d.astype(self.data.dtype)
        return (d * d if sqr_dists else d), i.astype(np.uint32)


def load_reference():
    pk = types.ModuleType("pykdtree")
    pkk = types.ModuleType("pykdtree.kdtree")
    pkk.KDTree = _CKDTreeStub
    pk.kdtree = pkk
    sys.modules["pykdtree"] = pk
    sys.modules["pykdtree.kdtree"] = pkk
    spec = importlib.util.spec_from_file_location("ref_astrolink_timed", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from astrolink_b200 import synth
    ref = load_reference()
    names = sys.argv[1:] or ["C3"]
    # JIT warm-up (both dtypes are float32 here)
    for d in (3, 6):
        w = ref.AstroLink(synth.two_gaussians(10000, d, 1, np.float32), workers=NCPU)
        w.run()
    out = {"machine_cores": NCPU, "kdtree": "scipy.spatial.cKDTree(workers=-1) stand-in for pykdtree (absent)",
           "reference": "william-h-oliver/astrolink, astrolink/astrolink.py imported in place, numba stages unmodified",
           "runs": {}}
    for name in names:
        wl = synth.WORKLOADS[name]
        P = wl["make"]()
        c = ref.AstroLink(P, workers=NCPU, **wl["kwargs"])
        t0 = time.perf_counter()
        c.run()
        total = time.perf_counter() - t0
        out["runs"][name] = {
            "workload": wl["desc"], "n": int(P.shape[0]), "d": int(P.shape[1]), "total_s": round(total, 3),
            "points_per_s": P.shape[0] / total,
            "stage_s": {"transform": round(c._transformTime, 3), "kNN_and_density": round(c._logRhoTime, 3),
                        "aggregate": round(c._aggregateTime, 3), "regression": round(c._regrTime, 3),
                        "rejection": round(c._rejTime, 3)},
            "clusters": int(c.clusters.shape[0]), "groups": int(c.groups.shape[0]),
        }
        print(f"{name}: {total:.2f} s  {P.shape[0] / total:.0f} points/s", file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
