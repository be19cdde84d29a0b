"""Builds astrolink_b200/libastrolink_b200.so from csrc/*.cu with nvcc for sm_100a.

The library is built IN-TREE (git-ignored, travels to the GPU box with the
repository snapshot).  The kNN kernel is instantiated once per (dtype, d) in
its own translation unit so the instantiations compile in parallel.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libastrolink_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-diag-suppress", "177,550",
]

PLAIN = ["tree.cu", "knn.cu", "stages.cu", "aggregate.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _units():
    units = []
    for f in PLAIN:
        if os.path.exists(os.path.join(CSRC, f)):
            units.append((f, [], f[:-3] + ".o"))
    for tname, t in (("f32", "float"), ("f64", "double")):
        for d in range(1, 17):
            units.append(("knn_inst.cu", [f"-DKNN_T={t}", f"-DKNN_TNAME={tname}", f"-DKNN_D={d}"], f"knn_{tname}_d{d}.o"))
    return units


def _deps_stamp():
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(tag, defs):
    """Development aid: builds libastrolink_b200_<tag>.so with extra -D flags (A/B tests via AL_LIB)."""
    global OBJ, LIB, NVCC_FLAGS
    saved = (OBJ, LIB, list(NVCC_FLAGS))
    OBJ = os.path.join(CSRC, "build_" + tag)
    LIB = os.path.join(HERE, f"libastrolink_b200_{tag}.so")
    NVCC_FLAGS = NVCC_FLAGS + list(defs)
    try:
        return build(force=False)
    finally:
        OBJ, LIB, NVCC_FLAGS = saved[0], saved[1], saved[2]


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _deps_stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    nvcc = _nvcc()
    units = _units()

    def compile_one(u):
        src, defs, obj = u
        cmd = [nvcc] + NVCC_FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {defs}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        objs = list(ex.map(compile_one, units))
    cmd = [nvcc, "-shared", "-o", LIB] + [os.path.join(OBJ, o) for o in objs] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
