"""Drift scores on a CPU-only box: the oracle restatement against the real scipy functions, and the numpy
emulation of the kernels' algorithm (tests/drift_walk.py mirrors csrc/drift_stats.cuh) against the oracle."""

import math

import numpy as np
import pytest
from scipy import stats


def _samples(rng, trial):
    m = int(rng.integers(20, 500))
    n = int(rng.integers(1, 250))
    if trial % 5 == 0:
        n = m  # equal sizes: scipy takes another formula (_compute_prob_outside_square), same probability
    if trial % 7 == 0:
        m, n = n + 5, m + 50  # batch larger than the reference
    if trial % 2:  # heavy ties, as the integer-valued credit columns have
        return np.sort(np.round(rng.normal(size=m) * 3) / 2), np.round(rng.normal(0.3, 1.2, size=n) * 3) / 2
    return np.sort(rng.normal(size=m)), rng.normal(0.2 * (trial % 3), 1, size=n)


def test_oracle_restatement_is_pinned_to_scipy():
    """oracle/drift.py's plain restatement of scipy's exact K-S == the compiled scipy functions, bit for bit."""
    from scipy.stats import _stats_pythran as sp

    from oracle import drift as od

    for m, n in [(50, 7), (30, 30), (64, 48), (300, 1), (300, 2), (1000, 37), (97, 100), (500, 125)]:
        g = math.gcd(m, n)
        lcm = m // g * n
        for h in sorted({1, 2, 3, lcm // 50 + 1, lcm // 10 + 1, lcm // 3 + 1, lcm - 1, lcm}):
            if 1 <= h <= lcm:
                want = min(max(sp._compute_outer_prob_inside_method(m, n, g, h), 0.0), 1.0)
                assert od.outer_prob_inside_method(m, n, g, h) == want
    rng = np.random.default_rng(0)
    for trial in range(40):
        ref, x = _samples(rng, trial)
        d, p = od.ks_2samp_exact(ref, x)
        r = stats.ks_2samp(ref, x, alternative="two-sided", method="exact")
        assert abs(d - r.statistic) < 1e-15
        assert abs(p - r.pvalue) <= 1e-12 * max(r.pvalue, 1e-300) + 1e-300
    for K in (2, 3, 7, 40):
        a = rng.integers(1, 5000, K)
        b = rng.integers(0, 30, K)
        b[0] += 1
        s, p = od.chi2_pvalue(a, b)
        r = stats.chi2_contingency(np.vstack((a, b)))
        assert abs(s - r[0]) <= 1e-12 * r[0] and abs(p - r[1]) <= 1e-12 * r[1]


def test_kernel_algorithm_matches_scipy():
    """The GPU formulation -- integer K-S numerator from two histograms over reference positions, anti-diagonal ring
    sweep of the lattice-path recursion, chi-squared tail by series / continued fraction -- against scipy."""
    import drift_walk as dw

    from oracle import drift as od

    rng = np.random.default_rng(1)
    for trial in range(60):
        ref, x = _samples(rng, trial)
        m, n = len(ref), len(x)
        num = dw.ks_numerator(ref, x)
        assert num == od.ks_numerator(ref, x)
        r = stats.ks_2samp(ref, x, alternative="two-sided", method="exact")
        assert abs(num / (m * n) - r.statistic) < 1e-15
        p, flag = dw.exact_p(m, n, num, check_bookkeeping=True)
        assert flag == 0 and abs(p - r.pvalue) <= 1e-12 * r.pvalue + 1e-300
    # single-row requests take a closed form instead of the sweep: both against scipy's recursion, every h
    from scipy.stats import _stats_pythran as sp

    for m in (1, 2, 3, 7, 30, 300):
        for h in range(1, m + 1):
            want = min(max(sp._compute_outer_prob_inside_method(m, 1, 1, h), 0.0), 1.0)
            assert abs(dw.exact_p(m, 1, h)[0] - want) <= 1e-13 * want
            assert abs(dw.exact_p(m, 1, h, force_ring=32)[0] - want) <= 1e-13 * want  # the sweep itself
    for h in (1, 2, 14999, 15000, 15001, 29999, 30000):
        want = min(max(sp._compute_outer_prob_inside_method(30000, 1, 1, h), 0.0), 1.0)
        assert abs(dw.exact_p(30000, 1, h)[0] - want) <= 1e-12 * want
    # a wider ring than needed changes nothing; one too narrow is refused (p underflows float32 there anyway)
    base = dw.exact_p(300, 40, 977)[0]
    assert dw.exact_p(300, 40, 977, force_ring=64)[0] == base == dw.exact_p(300, 40, 977, force_ring=256)[0]
    assert dw.exact_p(30000, 65536, 0)[0] == 1.0
    assert dw.exact_p(30000, 99991, 12345)[1] == 1  # lcm >= 2^31: scipy switches to the asymptotic formula
    for K in (1, 2, 3, 7, 40):
        a = rng.integers(1, 5000, K)
        b = rng.integers(0, 30, K)
        b[0] += 1
        if K >= 2:
            s, p = dw.chi2(a, b)
            r = stats.chi2_contingency(np.vstack((a, b)))
            assert abs(s - r[0]) <= 1e-12 * r[0] and abs(p - r[1]) <= 1e-11 * r[1]
        s, p = dw.chi2(a, b, [3, 1])  # two batch values outside the reference categories
        r = stats.chi2_contingency(np.vstack((np.concatenate((a, [0, 0])), np.concatenate((b, [3, 1])))))
        assert abs(s - r[0]) <= 1e-12 * r[0] and abs(p - r[1]) <= 1e-11 * r[1]


@pytest.mark.parametrize("n", [1, 37, 1000])
def test_kernel_algorithm_at_reference_size(curated, n):
    """m = 30 000 (the curated table) against request-sized batches: 30 000 + n anti-diagonals."""
    import drift_walk as dw

    from oracle import reference_pipeline as rp

    rng = np.random.default_rng(3)
    name = rp.NUMERIC_FEATURES[5]
    col = curated[name].to_numpy(float)
    ref = np.sort(col)
    x = col[rng.integers(0, len(col), n)] * (1.3 if n == 37 else 1.0)
    r = stats.ks_2samp(ref, x, alternative="two-sided", method="exact")
    num = dw.ks_numerator(ref, x)
    p, flag = dw.exact_p(len(ref), n, num)
    assert flag == 0 and abs(num / (len(ref) * n) - r.statistic) < 1e-15
    assert abs(p - r.pvalue) <= 1e-11 * r.pvalue


def test_drift_oracle_union_of_categories(curated):
    """alibi-detect counts over the union of reference and batch categories: an unseen value adds a (0, k) column."""
    from oracle import drift as od
    from oracle import reference_pipeline as rp

    ref = curated[rp.FEATURES]
    batch = ref.iloc[:50].copy()
    p0 = od.tabular_drift_p_values(ref, batch, rp.CATEGORICAL_FEATURES)
    batch.iloc[3, batch.columns.get_loc("sex")] = "unseen_value"
    p1 = od.tabular_drift_p_values(ref, batch, rp.CATEGORICAL_FEATURES)
    i = rp.FEATURES.index("sex")
    assert p1[i] < p0[i] and (np.delete(p1, i) == np.delete(p0, i)).all()
    assert p0.dtype == np.float32 and ((0 <= p0) & (p0 <= 1)).all()
    scores = od.drift_scores(ref, batch, rp.CATEGORICAL_FEATURES)
    assert len(scores) == 23 and all(0.0 <= v <= 1.0 for v in scores)


def test_frozen_detector_outputs_are_reproduced(curated, inference, iforest):
    """tests/golden/expected_detectors.npz (library outputs frozen by make_golden_detectors.py) == what the oracle
    computes on this box: pins the scipy / sklearn behaviour the GPU tests compare against."""
    import os

    import scipy
    import sklearn

    from oracle import datasets
    from oracle import drift as od
    from oracle import reference_pipeline as rp

    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden_detectors", os.path.join(sys_path, "make_golden_detectors.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    exp = datasets.load_expected("detectors")
    ref = curated[rp.FEATURES]
    for key, batch in mg.drift_batches(curated, inference).items():
        stat, p = mg.drift_reference_values(ref, batch)
        assert np.abs(stat - exp[f"drift_stat_{key}"]).max() <= 1e-12 * np.abs(stat).max()
        assert (np.abs(p - exp[f"drift_p_{key}"]) <= 1e-10 * exp[f"drift_p_{key}"] + 1e-300).all()
        p32 = od.tabular_drift_p_values(ref, batch, rp.CATEGORICAL_FEATURES)
        assert (p32 == p.astype(np.float32)).all()
    if str(exp["sklearn_version"]) == sklearn.__version__:  # the isolation trees depend on the library's RNG stream
        got = -iforest.decision_function(curated[rp.NUMERIC_FEATURES].iloc[:3000].to_numpy())
        assert np.abs(got - exp["iforest_score_head3000"]).max() <= 1e-12
    assert str(exp["scipy_version"]) and scipy.__version__


def test_row_scan_form_matches_scipy():
    """The row-scan form of the exact p-value (what k_drift_finish runs for request-sized batches: n prefix sums over the
    unnormalised path counts, binomials outside the band, rows scaled by powers of two) against scipy's compiled
    recursion, small lattices exhaustively in h and the reference-table size for a few batch sizes."""
    import math

    import drift_walk as dw
    from scipy.stats import _stats_pythran as sp

    worst = 0.0
    for m, n in [(50, 7), (64, 48), (300, 2), (300, 16), (1000, 37), (500, 125), (97, 100), (31, 31)]:
        g = math.gcd(m, n)
        lcm = m // g * n
        for h in sorted({1, 2, 3, lcm // 50 + 1, lcm // 10 + 1, lcm // 3 + 1, lcm - 1, lcm}):
            if 1 <= h <= lcm:
                want = min(max(sp._compute_outer_prob_inside_method(max(m, n), min(m, n), g, h), 0.0), 1.0)
                got, flag = dw.exact_p_rows(m, n, h * g)
                assert flag == 0
                worst = max(worst, abs(got - want) / max(want, 1e-300))
    for n, d in ((2, 0.7), (16, 0.33), (100, 0.12), (128, 0.09)):
        m = 30000
        g = math.gcd(m, n)
        h = max(1, int(d * (m // g) * n))
        want = min(max(sp._compute_outer_prob_inside_method(m, n, g, h), 0.0), 1.0)
        got, flag = dw.exact_p_rows(m, n, h * g)
        worst = max(worst, abs(got - want) / max(want, 1e-300))
    assert worst <= 1e-12, worst
    # and equal to the anti-diagonal sweep of the same kernel
    for m, n, num in ((30000, 16, 30000 * 16 // 3), (30000, 128, 30000 * 128 // 9), (4096, 5, 4096 * 5 // 2)):
        a, _ = dw.exact_p(m, n, num)
        b, _ = dw.exact_p_rows(m, n, num)
        assert abs(a - b) <= 1e-12 * max(a, 1e-300)


def test_row_scan_ring_schedule_matches_scipy():
    """The shared-memory schedule of the row scan (rows_scan_smem: ring of `cap` slots updated in place, L-cell thread shares,
    cells entering the band written into the ring first) emulated step by step: a slot read before it is written, or
    overwritten while still needed, would poison the result.  Small rings force many wrap-arounds."""
    import math

    import drift_walk as dw
    from scipy.stats import _stats_pythran as sp

    worst, ran = 0.0, 0
    for m, n, cap, nt in [(300, 16, 64, 8), (300, 16, 301, 32), (1000, 37, 128, 16), (500, 125, 40, 4), (2048, 3, 1500, 64), (997, 64, 97, 8)]:
        g = math.gcd(m, n)
        lcm = m // g * n
        for h in sorted({1, 2, 3, lcm // 50 + 1, lcm // 20 + 1, lcm // 10 + 1, lcm // 5 + 1, lcm // 3 + 1}):
            res = dw.exact_p_rows_ring(m, n, h * g, cap=cap, nt=nt)
            if res is None:
                continue  # band wider than the ring: the kernel takes the global-scratch form or the sweep
            want = min(max(sp._compute_outer_prob_inside_method(max(m, n), min(m, n), g, h), 0.0), 1.0)
            worst = max(worst, abs(res[0] - want) / max(want, 1e-300))
            ran += 1
    assert ran >= 20 and worst <= 1e-12, (ran, worst)
    m, n = 30000, 100
    h = int(0.12 * m * n / math.gcd(m, n))
    want = min(max(sp._compute_outer_prob_inside_method(m, n, math.gcd(m, n), h), 0.0), 1.0)
    got, _ = dw.exact_p_rows_ring(m, n, h * math.gcd(m, n))
    assert abs(got - want) <= 1e-12 * max(want, 1e-300)


def test_native_kstwo_sf_matches_scipy():
    """b2f_kstwo_sf (host, csrc/drift_api.cuh) restates the branch scipy's ks_2samp takes when the exact method is not
    applicable (lcm of the sample sizes >= 2^31): kstwo.sf(D, round(m n / (m + n))).  Checked against scipy over the
    whole range of D for the effective sample sizes that reach that branch with a 30 000-row reference table."""
    from scipy.stats import distributions

    from databricks_kubernetes_mlops_poc_b200 import _cabi

    lib = _cabi.load_library()
    for n in (5000, 21145, 25000, 29999):
        xs = list(np.geomspace(1e-5, 0.9, 60)) + [1 / n, 0.6 / n, 0.4 / n, 1 - 0.5 / n, 0.5, 0.51, 0.0, 1.0, 1.5, -0.1]
        for x in xs:
            want = float(distributions.kstwo.sf(x, n))
            got = lib.b2f_kstwo_sf(float(x), float(n))
            assert abs(got - want) <= 1e-10 * max(want, 1e-300) + 1e-300, (n, x, got, want)
    assert np.isnan(lib.b2f_kstwo_sf(float("nan"), 100.0))
