"""Host-side statistics of AstroLink: the prominence model fit and the cluster
extraction.  These stay in Python/NumPy/SciPy (north-star item 5); they consume
the small `groups` / `prominences` arrays the GPU aggregation produces.

Restated from the behaviour of the reference's
``compute_significances`` (astrolink/astrolink.py:828-907) and
``extract_clusters`` (astrolink/astrolink.py:909-980); pinned against the
reference's outputs by tests/test_stats.py (golden vectors).
"""
import numpy as np
from scipy.optimize import minimize
from scipy.special import betaln
from scipy.stats import beta, norm


def _initial_beta_guess(p_sorted):
    """Starting point of the fit (astrolink.py:866-877): the mean of a
    method-of-moments estimate and a mode-based estimate, the mode taken from a
    Freedman-Diaconis histogram of the sorted prominences."""
    N = p_sorted.size
    iqr = p_sorted[3 * N // 4] - p_sorted[N // 4]
    binwidth = 2.0 * iqr * N ** (-1.0 / 3.0)
    # (floor of the quotient, as numba's fastmath `//` in the reference; np.floor_divide is ten times slower on floats)
    counts = np.bincount(np.floor(p_sorted / binwidth).astype(np.int64))
    mode = binwidth * (np.argmax(counts) + 0.5)
    mu = p_sorted.mean()
    var = p_sorted.var()
    one_minus_mu = 1.0 - mu
    common = mu * one_minus_mu / var - 1.0
    a_mom, b_mom = mu * common, one_minus_mu * common
    a_mode = (2.0 * mode - 1.0) * mu / (one_minus_mu * mode - (1.0 - mode) * mu)
    b_mode = a_mode * one_minus_mu / mu
    return np.array([0.5 * (a_mom + a_mode), 0.5 * (b_mom + b_mode)])


class _W1Objective:
    """Wasserstein-1 distance between the numerical cdf of Beta(a, b), integrated
    with the midpoint rule over the gaps between consecutive sorted prominences,
    and the empirical cdf (astrolink.py:880-885, 896-907).

    Evaluated in the log domain (one exp per midpoint instead of two powers);
    the minimisation itself follows the reference exactly (L-BFGS-B with bounds,
    `jac='3-point'`) so that pFit lands on the same point of this non-smooth
    objective."""

    def __init__(self, p_sorted):
        N = p_sorted.size
        edges = np.concatenate((np.zeros(1), p_sorted))
        self.widths = edges[1:] - edges[:-1]
        mid = 0.5 * (edges[1:] + edges[:-1])
        # a midpoint of exactly 0 (or 1) only occurs in a zero-width gap, whose mass is 0 anyway
        safe = self.widths > 0
        self.log_mid = np.where(safe, np.log(np.where(safe, mid, 0.5)), 0.0)
        self.log_omm = np.where(safe, np.log1p(-np.where(safe, mid, 0.5)), 0.0)
        half = 1.0 / (2 * N)
        self.quantiles = np.linspace(half, 1.0 - half, N)

    def __call__(self, p):
        logpdf = (p[0] - 1.0) * self.log_mid + (p[1] - 1.0) * self.log_omm - betaln(p[0], p[1])
        cdf = np.cumsum(np.exp(logpdf) * self.widths)
        return float(np.sum(np.abs(cdf - self.quantiles) * self.widths))


def fit_prominence_model(prom_leq, objective=None, p_sorted=None):
    """Returns (pFit [2], success).  astrolink.py:848-853.  `objective` / `p_sorted` let the caller
    supply an evaluator of the same function (the GPU one of al_w1_eval) and the sorted values."""
    if p_sorted is None:
        p_sorted = np.sort(np.asarray(prom_leq, dtype=np.float64))
    N = p_sorted.size
    guess = _initial_beta_guess(p_sorted)
    tol = 10.0 ** (-np.ceil(np.log10(N)) - 3)
    if objective is None:
        objective = _W1Objective(p_sorted)
    sol = minimize(objective, guess, jac="3-point", bounds=((1, np.inf), (1, np.inf)), tol=tol)
    if sol.success:
        return np.asarray(sol.x, dtype=np.float64), True
    return guess, False


def significances_of(prominences, pFit):
    """groups_sigs = norm.isf(beta.sf(prominences, a, b)) on both columns (astrolink.py:858)."""
    return norm.isf(beta.sf(prominences, pFit[0], pFit[1]))


def auto_threshold(n_groups):
    """S = 'auto' (astrolink.py:934)."""
    return norm.isf(1 - 0.5 ** (1 / n_groups))


def extract(groups, groups_sigs, S, h_style, n_samples):
    """Clusters, hierarchical ids and significances (astrolink.py:935-978).

    A group's smaller side [leq_start, leq_end) is a cluster when its
    significance reaches S.  With h_style == 1 the larger side
    [geq_start, leq_start) of such a row is a cluster too when it is itself
    significant; of the rows sharing a geq_start only the first is kept.  Rows
    are ordered by (start ascending, end descending, insertion), the root
    [0, n] is prepended, and ids are assigned by walking the nesting."""
    sel = groups_sigs[:, 1] >= S
    clusters = groups[sel, 1:]
    sigs = groups_sigs[sel, 1]
    if h_style == 1:
        both = np.logical_and(sel, groups_sigs[:, 0] >= S)
        geq = groups[both, :2]
        geq_sigs = groups_sigs[both, 0]
        _, first_rows = np.unique(geq[:, 0], return_index=True)
        keep = np.zeros(geq.shape[0], dtype=np.bool_)
        keep[first_rows] = True
        clusters = np.vstack((clusters, geq[keep]))
        sigs = np.concatenate((sigs, geq_sigs[keep]))
        order = np.lexsort((np.arange(clusters.shape[0]), n_samples - clusters[:, 1], clusters[:, 0]))
        clusters = clusters[order]
        sigs = sigs[order]
    sigs = np.concatenate((np.array([np.inf]), sigs))
    clusters = np.vstack((np.array([[0, n_samples]], dtype=np.uint32), clusters))
    ids = ["1"]
    n_children = np.zeros(clusters.shape[0], dtype=np.uint32)
    open_parents = [0]
    for i in range(1, clusters.shape[0]):
        start = clusters[i, 0]
        while start >= clusters[open_parents[-1], 1]:
            open_parents.pop()
        parent = open_parents[-1]
        open_parents.append(i)
        n_children[parent] += 1
        ids.append(f"{ids[parent]}-{n_children[parent]}")
    return clusters, np.array(ids), sigs
