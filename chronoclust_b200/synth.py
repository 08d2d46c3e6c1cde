"""Synthetic cytometry-shaped time series.

gen(N, D, T, C, seed) -> list of T arrays [N, D] fp64 in [0, 1], C-contiguous.  C Gaussian
populations (sigma 0.025 per marker, the spread of the reference's own synthetic_dataset in scaled
units) with Dirichlet(2) proportions, 1 % uniform background noise, centres drifting by N(0, 0.01)
between timepoints.  Deterministic for a given seed; used with normalise_data=False.
"""
import numpy as np

CONFIGS = {
    # name: (N, D, T, C, seed, epsilon, pi)           -- BASELINE.json configs[1], configs[2]
    "C2": (1_000_000, 12, 5, 20, 1234, 0.05, 12),
    "C3": (2_000_000, 40, 5, 40, 1234, 0.10, 40),
}
# sample_run.py:6-16 of the reference, used unless a run states otherwise
BASE_PARAMS = dict(beta=0.2, delta=0.05, lambda_=2, k=4, mu=0.01, omicron=0.00000435, upsilon=6.5)


def gen(N, D, T, C, seed, sigma=0.025, noise=0.01, drift=0.01, aniso=None):
    """aniso=(fraction, factor): that fraction of each population's markers gets sigma*factor
    (used by the parity stress cases to make the pi < D feasibility gate bite)."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.15, 0.85, size=(C, D))
    prop = rng.dirichlet(2.0 * np.ones(C))
    sig = np.full((C, D), sigma)
    if aniso is not None:
        arng = np.random.default_rng(seed + 7919)
        sig = np.where(arng.random((C, D)) < aniso[0], sigma * aniso[1], sigma)
    out = []
    for _ in range(T):
        lab = rng.choice(C, size=N, p=prop)
        x = centres[lab] + rng.normal(0.0, 1.0, size=(N, D)) * sig[lab]
        nz = rng.random(N) < noise
        x[nz] = rng.random((int(nz.sum()), D))
        np.clip(x, 0.0, 1.0, out=x)
        out.append(np.ascontiguousarray(x, dtype=np.float64))
        centres = np.clip(centres + rng.normal(0.0, drift, size=(C, D)), 0.05, 0.95)
    return out


def config_params(name):
    """Reference-style config dict for a named BASELINE config."""
    N, D, T, C, seed, eps, pi = CONFIGS[name]
    p = dict(BASE_PARAMS)
    return {"beta": p["beta"], "delta": p["delta"], "epsilon": eps, "lambda": p["lambda_"], "k": p["k"],
            "mu": p["mu"], "pi": pi, "omicron": p["omicron"], "upsilon": p["upsilon"]}


def gen_offline_stress(M, D=40, seed=1, ncentres=50, spread=0.03):
    """Config C4: M pcore MCs as (centroid, weight, core flag) for the offline stress."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(0.1, 0.9, size=(ncentres, D))
    lab = rng.integers(0, ncentres, size=M)
    cen = ctr[lab] + rng.normal(0.0, spread, size=(M, D))
    w = rng.integers(5, 200, size=M).astype(np.float64)
    return np.ascontiguousarray(cen), w, (w >= 20)
