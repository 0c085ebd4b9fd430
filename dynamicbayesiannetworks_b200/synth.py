"""Synthetic piecewise-stationary sparse VAR(1) series for the BASELINE shapes.

The reference ships its own generator (`/root/reference/src/dyban/generateTestData.py:16-157`,
seed 42, coefficients from `coefs.txt`) that only produces the small simulated-data example.
SURVEY.md section 8(d) fixes the recipe used for BASELINE configs 4 and 5 instead: a sparse VAR(1)
per regime, 1..max_fan_in parents per gene, spectral radius 0.8, unit-variance noise, z-scored columns.
"""
import numpy as np


def make_synthetic(n, T, n_changepoints=3, max_fan_in=4, seed=20260119, dtype=np.float64):
    """Return (data[T, n], true_cps, B) with `n_changepoints` equally spaced regime switches.

    data  : z-scored time series, rows are time points (what `Network([data], ...)` takes).
    true_cps : 0-based time indices where the regime switches.
    B     : list of per-regime coefficient matrices, B[h][child, parent].
    """
    rng = np.random.default_rng(seed)
    n_reg = n_changepoints + 1
    cps = [int(round(T * (h + 1) / n_reg)) for h in range(n_changepoints)]
    mats = []
    for _ in range(n_reg):
        B = np.zeros((n, n))
        for child in range(n):
            fan = int(rng.integers(1, max_fan_in + 1))
            fan = min(fan, n - 1) if n > 1 else 0
            if fan == 0:
                continue
            cand = np.delete(np.arange(n), child)
            parents = rng.choice(cand, size=fan, replace=False)
            mag = rng.uniform(0.3, 0.8, size=fan)
            sgn = rng.choice([-1.0, 1.0], size=fan)
            B[child, parents] = mag * sgn
        rho = np.max(np.abs(np.linalg.eigvals(B))) if n > 1 else 0.0
        if rho > 0:
            B *= 0.8 / rho
        mats.append(B)
    x = np.zeros((T, n))
    x[0] = rng.standard_normal(n)
    bounds = [0] + cps + [T]
    reg = 0
    for t in range(T - 1):
        while t + 1 >= bounds[reg + 1] and reg + 1 < n_reg:
            reg += 1
        x[t + 1] = mats[reg] @ x[t] + rng.standard_normal(n)
    x = (x - x.mean(axis=0)) / x.std(axis=0)
    return x.astype(dtype), cps, mats
