"""Host statistics (astrolink_b200/stats.py) against the reference's
compute_significances / extract_clusters outputs in the golden vectors."""
import numpy as np
import pytest

from astrolink_b200 import stats
from tests.conftest import golden_names


def _with_stats(golden):
    return [n for n in golden_names() if "pFit" in golden(n)]


def test_some_cases_have_stats(golden):
    assert len(_with_stats(golden)) >= 10


@pytest.mark.parametrize("name", golden_names())
def test_fit_and_significances(name, golden):
    g = golden(name)
    if "pFit" not in g:
        pytest.skip("too few groups for a fit in this case")
    pFit, ok = stats.fit_prominence_model(g["prominences"][:, 1])
    assert ok
    assert np.allclose(pFit, g["pFit"], rtol=2e-3), (pFit, g["pFit"])
    sig_ref = stats.significances_of(g["prominences"], g["pFit"])
    fin = np.isfinite(g["groups_sigs"])
    assert np.allclose(sig_ref[fin], g["groups_sigs"][fin], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", golden_names())
def test_extract_clusters_exact(name, golden):
    g = golden(name)
    if "pFit" not in g:
        pytest.skip("no clusters golden")
    h_style = int(g["kw_h_style"]) if "kw_h_style" in g else 1
    n = g["P"].shape[0]
    S = float(g["kw_S"]) if "kw_S" in g else stats.auto_threshold(g["prominences"].shape[0])
    assert np.isclose(S, float(g["S"]))
    clusters, ids, sigs = stats.extract(g["groups"], g["groups_sigs"], S, h_style, n)
    assert clusters.dtype == np.uint32 and (clusters == g["clusters"]).all()
    assert list(ids) == list(g["ids"])
    assert np.allclose(sigs, g["significances"])
