"""utils/metrics.py:24-77 restated without scikit-learn; checked against closed forms and against scikit-learn's
pairwise kernels (the implementation upstream calls)."""
import numpy as np
import pytest

from smd_b200 import metrics


def test_frechet_distance_closed_forms():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((4000, 6))
    assert abs(metrics.frechet_distance(a, a)) < 1e-8
    # shifting one set by d adds |d|^2; scaling by s (same shape) gives sum_i (1 - s)^2 var_i
    d = np.array([0.5, -1.0, 0.0, 2.0, 0.0, 0.25])
    assert abs(metrics.frechet_distance(a, a + d) - float(d @ d)) < 1e-6
    var = np.cov(a, rowvar=False)
    mu = a.mean(0)
    expect = float(0.25 * (mu @ mu) + np.trace(var) * (1 - 0.5) ** 2)
    assert abs(metrics.frechet_distance(a, 0.5 * a) - expect) < 1e-6


def test_mmd_matches_sklearn_kernels():
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal((60, 5)), 0.7 * rng.standard_normal((45, 5)) + 0.2
    rbf = lambda a, b: sk.pairwise.rbf_kernel(a, b, 0.3)
    want = rbf(x, x).mean() + rbf(y, y).mean() - 2 * rbf(x, y).mean()
    assert abs(metrics.mmd_rbf(x, y, gamma=0.3) - want) < 1e-12
    pol = lambda a, b: sk.pairwise.polynomial_kernel(a, b, 3, 0.5, 1.0)
    want = pol(x, x).mean() + pol(y, y).mean() - 2 * pol(x, y).mean()
    assert abs(metrics.mmd_polynomial(x, y, degree=3, gamma=0.5, coef0=1.0) - want) < 1e-9 * max(1.0, abs(want))
    assert metrics.mmd_rbf(x, x) < 1e-12 and metrics.mmd_rbf(x, y) > 0
