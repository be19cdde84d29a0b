"""The C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "astrolink_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(al_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from astrolink_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/astrolink_b200.h but not exported"
    lib.al_version.restype = ctypes.c_int
    assert lib.al_version() == 1


def test_binding_table_matches_header():
    from astrolink_b200 import _native
    assert sorted(_native.exported_symbols()) == _declared()


def test_workspace_queries_run_without_gpu():
    from astrolink_b200 import _native
    L = _native.lib()
    assert L.al_index_bytes(0, 10**6, 6) > 10**6 * 24
    assert L.al_aggregate_workspace_bytes(0, 10**6, 6 * 10**6) > 0
    assert L.al_index_positions(33) == 64


def test_constructor_asserts_like_the_reference_and_fails_loudly_without_gpu():
    import numpy as np
    import torch
    from astrolink_b200 import AstroLink, _native
    P = np.zeros((100, 2))
    with pytest.raises(AssertionError, match="needs to be a 2D numpy array"):
        AstroLink([1, 2, 3])
    with pytest.raises(AssertionError, match="'k_den' must be an integer less than"):
        AstroLink(P, k_den=101)
    with pytest.raises(AssertionError, match="'adaptive' must be set as either"):
        AstroLink(P, adaptive=2)
    with pytest.raises(AssertionError, match="'k_link' needs to be an integer"):
        AstroLink(P, k_link=1)
    with pytest.raises(AssertionError, match="'h_style' must be set"):
        AstroLink(P, h_style=3)
    with pytest.raises(AssertionError, match="'S' must be set to 'auto'"):
        AstroLink(P, S=40)
    with pytest.raises(AssertionError, match="'d_intrinsic' must be None"):
        AstroLink(P, d_intrinsic=3)
    with pytest.raises(AssertionError, match="'weights' must be None"):
        AstroLink(P, weights=np.ones(5))
    if not torch.cuda.is_available():
        with pytest.raises(_native.NativeError, match="no CPU fallback"):
            AstroLink(P)
