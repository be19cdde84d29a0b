"""Seeded synthetic point clouds for the BASELINE.json configs.

There is no network and the reference's one data file (data/newhalo_young.npy,
config C2) is an absent LFS blob, so every workload is generated here from a
fixed seed.  All generators return C-contiguous arrays.
"""
import numpy as np


def readme_example(seed=0):
    """C1: the README recipe (/root/reference/README.md:44-52): 1000 uniform
    background points, 2000 two-moons points, two 500-point Gaussians; float64."""
    from sklearn import datasets as data
    np.random.seed(seed)
    background = np.random.uniform(-2, 2, (1000, 2))
    moons, _ = data.make_moons(n_samples=2000, noise=0.1)
    moons -= np.array([[0.5, 0.25]])
    gauss_1 = np.random.normal(-1.25, 0.2, (500, 2))
    gauss_2 = np.random.normal(1.25, 0.2, (500, 2))
    return np.ascontiguousarray(np.vstack([background, moons, gauss_1, gauss_2]))


def two_gaussians(n_each, d, seed, dtype=np.float64):
    """The reference's own test recipe (tests/test_astrolink.py:7-9,18-20), seeded."""
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 1, (n_each, d))
    centre = np.zeros(d)
    centre[0] = 10
    b = rng.normal(centre, 1, (n_each, d))
    return np.ascontiguousarray(np.concatenate((a, b), axis=0).astype(dtype))


def halo6d(n, seed=2, n_sub=64, sub_frac=0.5, nested=False, dtype=np.float32):
    """C2 / C4: 6-D phase-space halo (x,y,z kpc; vx,vy,vz km/s): a smooth
    Gaussian host plus `n_sub` subhaloes with power-law masses; optionally each
    subhalo holds up to 3 sub-subhaloes (C4)."""
    rng = np.random.default_rng(seed)
    n_host = int(n * (1 - sub_frac))
    host_sig = np.array([30, 30, 30, 150, 150, 150], dtype=np.float64)
    parts = [rng.normal(0, 1, (n_host, 6)) * host_sig]
    n_rem = n - n_host
    m = rng.pareto(0.9, n_sub) + 1.0          # dN/dm ~ m^-1.9
    m = m / m.sum()
    counts = np.floor(m * n_rem).astype(np.int64)
    counts[np.argmax(counts)] += n_rem - counts.sum()
    centres = rng.normal(0, 1, (n_sub, 6)) * np.array([25, 25, 25, 120, 120, 120.0])
    scales = rng.uniform(0.3, 3.0, n_sub)
    for c, ctr, s in zip(counts, centres, scales):
        if c <= 0:
            continue
        sig = s * np.array([1, 1, 1, 5, 5, 5.0])
        if nested and c > 2000:
            k = int(rng.integers(0, 4))
            c_sub = int(0.15 * c) if k else 0
            parts.append(ctr + rng.normal(0, 1, (c - c_sub * k, 6)) * sig)
            for _ in range(k):
                sctr = ctr + rng.normal(0, 1, 6) * sig
                parts.append(sctr + rng.normal(0, 1, (c_sub, 6)) * sig * 0.15)
        else:
            parts.append(ctr + rng.normal(0, 1, (c, 6)) * sig)
    P = np.concatenate(parts, axis=0)
    P = P[rng.permutation(P.shape[0])]
    return np.ascontiguousarray(P[:n].astype(dtype))


def gmm3d_nested(n, seed=3, dtype=np.float32):
    """C3: 3-level nested Gaussian mixture in 3-D: 8 parents (sigma 1), 4
    children each (sigma .25), 2 grandchildren per child (sigma .06); mass split
    40/35/15 % plus 10 % uniform background on the bounding box."""
    rng = np.random.default_rng(seed)
    n_par, n_chi, n_gra = int(0.40 * n), int(0.35 * n), int(0.15 * n)
    n_bg = n - n_par - n_chi - n_gra
    pc = rng.uniform(-10, 10, (8, 3))
    cc = (pc[:, None, :] + rng.normal(0, 1.0, (8, 4, 3))).reshape(-1, 3)
    gc = (cc[:, None, :] + rng.normal(0, 0.25, (32, 2, 3))).reshape(-1, 3)
    P1 = pc[rng.integers(0, 8, n_par)] + rng.normal(0, 1.0, (n_par, 3))
    P2 = cc[rng.integers(0, 32, n_chi)] + rng.normal(0, 0.25, (n_chi, 3))
    P3 = gc[rng.integers(0, 64, n_gra)] + rng.normal(0, 0.06, (n_gra, 3))
    fg = np.concatenate((P1, P2, P3), axis=0)
    bg = rng.uniform(fg.min(axis=0), fg.max(axis=0), (n_bg, 3))
    P = np.concatenate((fg, bg), axis=0)
    P = P[rng.permutation(n)]
    return np.ascontiguousarray(P.astype(dtype))


def cosmo3d(n, seed=5, dtype=np.float32):
    """C5: 30 % uniform in [0,1000]^3 + 70 % in Gaussian clumps with
    log-uniform sizes (0.5-20)."""
    rng = np.random.default_rng(seed)
    n_bg = int(0.3 * n)
    n_cl = n - n_bg
    n_clumps = max(8, min(20000, n // 3500))
    ctr = rng.uniform(0, 1000, (n_clumps, 3))
    sig = np.exp(rng.uniform(np.log(0.5), np.log(20), n_clumps))
    which = rng.integers(0, n_clumps, n_cl)
    cl = ctr[which] + rng.normal(0, 1, (n_cl, 3)) * sig[which, None]
    bg = rng.uniform(0, 1000, (n_bg, 3))
    P = np.concatenate((cl, bg), axis=0)
    P = P[rng.permutation(n)]
    return np.ascontiguousarray(P.astype(dtype))


WORKLOADS = {
    "C1": dict(desc="README example 4000x2 f64", make=lambda: readme_example(0), kwargs={}),
    "C2": dict(desc="synthetic stand-in for newhalo_young 200359x6 f32", make=lambda: halo6d(200359, seed=2), kwargs={}),
    "C3": dict(desc="1M x 3D nested GMM f32", make=lambda: gmm3d_nested(10**6, seed=3), kwargs=dict(adaptive=0)),
    "C4": dict(desc="10M x 6D synthetic phase-space halo f32", make=lambda: halo6d(10**7, seed=4, n_sub=256, nested=True), kwargs={}),
    "C5": dict(desc="100M x 3D cosmological-style cloud f32", make=lambda: cosmo3d(10**8, seed=5), kwargs=dict(adaptive=0)),
}
