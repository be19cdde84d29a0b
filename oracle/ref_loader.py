"""Loads the reference implementation IN PLACE from /root/reference (never copied).

Only usable in the build container (the GPU box has no /root/reference); used
by oracle/gen_golden.py to pin the oracle and to generate tests/golden/*.npz.

The reference imports ``pykdtree.kdtree.KDTree`` (astrolink.py:16), which is not
installed here and not installable offline.  A stub module with the same call
surface (``KDTree(data).query(q, k=..., sqr_dists=True)`` -> (d2, idx uint32),
astrolink.py:270,279) is injected; the stub answers with the oracle's exact kNN
(defined fma distance chain, index tie-break), so every downstream reference
stage consumes neighbour lists that obey the contract the CUDA path is tested
against.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_FILE = "/root/reference/astrolink/astrolink.py"


def available():
    return os.path.exists(REF_FILE)


class _StubKDTree:
    def __init__(self, data, leafsize=16):
        self.data = np.ascontiguousarray(data)

    def query(self, q, k=1, sqr_dists=False, **kw):
        from . import oracle
        q = np.ascontiguousarray(q)
        assert q.shape == self.data.shape and (q == self.data).all(), "stub supports self-queries only"
        d2, idx = oracle.knn(self.data, k)
        return (d2 if sqr_dists else np.sqrt(d2)), idx


def load():
    if "ref_astrolink" in sys.modules:
        return sys.modules["ref_astrolink"]
    pk = types.ModuleType("pykdtree")
    pkk = types.ModuleType("pykdtree.kdtree")
    pkk.KDTree = _StubKDTree
    pk.kdtree = pkk
    sys.modules.setdefault("pykdtree", pk)
    sys.modules.setdefault("pykdtree.kdtree", pkk)
    spec = importlib.util.spec_from_file_location("ref_astrolink", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_astrolink"] = mod
    spec.loader.exec_module(mod)
    return mod
