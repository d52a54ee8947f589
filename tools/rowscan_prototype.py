"""Prototype (numpy, CPU) of the row-wise formulation of the exact two-sample K-S p-value -- groundwork for round 2.

The shipped kernel (csrc/drift_stats.cuh) sweeps anti-diagonals: m + n dependent steps.  For request-sized batches
(2 <= n <~ 250 against m = 30 000 reference rows) the same probabilities can be produced row by row: for fixed j

    P(i, j) = a_i * P(i-1, j) + b_i,     a_i = i / (i + j),   b_i = j / (i + j) * P(i, j-1)      inside the band
            = 1  (a_i = 0, b_i = 1)                                                                outside it

is a first-order affine recurrence in i, i.e. an inclusive SCAN over i of the maps x -> a_i x + b_i under composition
(a2, b2) o (a1, b1) = (a2 a1, a2 b1 + b2): n scans of length m, each log-depth, instead of m + n dependent steps.
This file checks the formulation and its rounding behaviour (a Hillis-Steele scan evaluates the products in a very
different order from scipy's sequential loop) against scipy's compiled recursion.

    python tools/rowscan_prototype.py
"""
import math
import time

import numpy as np
from scipy.stats import _stats_pythran as sp


def affine_scan(a, b):
    """Inclusive scan of x -> a_i x + b_i under composition, Hillis-Steele order (what a block-wide GPU scan does)."""
    a, b = a.copy(), b.copy()
    d = 1
    n = len(a)
    while d < n:
        a2, b2 = a.copy(), b.copy()
        # element i composes (its own map) after (the map accumulated at i - d)
        b2[d:] = a[d:] * b[:-d] + b[d:]
        a2[d:] = a[d:] * a[:-d]
        a, b = a2, b2
        d *= 2
    return a, b


def outer_prob_rowscan(m, n, g, h):
    if m < n:
        m, n = n, m
    mg, ng = m // g, n // g
    i = np.arange(m + 1, dtype=np.int64)
    prev = None
    for j in range(n + 1):
        inband = np.abs(ng * i - mg * j) < h
        a = np.where(inband, i / np.maximum(i + j, 1), 0.0)
        if j == 0:
            b = np.where(inband, 0.0, 1.0)
        else:
            b = np.where(inband, j / np.maximum(i + j, 1) * prev, 1.0)
        # i = 0: P(0, j) = 0 inside the band, 1 outside: a_0 = 0, b_0 = that value
        a[0] = 0.0
        b[0] = 0.0 if inband[0] else 1.0
        _, x = affine_scan(a, b)  # x_i = composite applied to anything (a_0 = 0 kills the seed)
        prev = x
    return float(min(max(prev[m], 0.0), 1.0))


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    worst = 0.0
    for m, n in [(50, 7), (64, 48), (300, 2), (300, 16), (1000, 37), (500, 125), (97, 100)]:
        g = math.gcd(m, n)
        lcm = m // g * n
        for h in sorted({1, 2, 3, lcm // 50 + 1, lcm // 10 + 1, lcm // 3 + 1, lcm - 1, lcm}):
            if 1 <= h <= lcm:
                want = min(max(sp._compute_outer_prob_inside_method(m, n, g, h), 0.0), 1.0)
                got = outer_prob_rowscan(m, n, g, h)
                rel = abs(got - want) / max(want, 1e-300)
                worst = max(worst, rel)
    print("small cases: worst relative difference to scipy", worst)
    for n, d in ((2, 0.7), (16, 0.33), (100, 0.12), (250, 0.07)):
        m = 30000
        g = math.gcd(m, n)
        h = max(1, int(d * (m // g) * n))
        t0 = time.time()
        want = min(max(sp._compute_outer_prob_inside_method(m, n, g, h), 0.0), 1.0)
        t1 = time.time()
        got = outer_prob_rowscan(m, n, g, h)
        print(f"m=30000 n={n:4d} h={h}: scipy {want:.15e} rowscan {got:.15e} rel {abs(got - want) / max(want, 1e-300):.2e} "
              f"(scipy {1e3 * (t1 - t0):.1f} ms; {n + 1} scans of {m + 1} elements, 15 Hillis-Steele levels each)")
