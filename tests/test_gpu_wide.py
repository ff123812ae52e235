"""Integrands with more components than the reduce kernel is instantiated for (``_lib.MAX_REDUCE_NF`` = 8):
``Integrator._reduce_wide`` runs the kernel on column subsets.  The reference has no such limit
(pyx:2136-2197 loops over ``fcn.size``); here the wide result must equal, block by block, what narrow
integrands give on the same samples."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _components(x):
    c = [x[:, 0], x[:, 1] ** 2, x[:, 0] * x[:, 2], np.cos(x[:, 1]), np.exp(-x[:, 2]), x[:, 0] ** 3, x[:, 1] * x[:, 2],
         np.sin(3 * x[:, 0]), x[:, 2] ** 2, x[:, 0] + x[:, 1], x[:, 1] ** 3, np.ones(x.shape[0]), x[:, 0] * x[:, 1] * x[:, 2]]
    return np.stack(c, axis=1)


@pytest.mark.parametrize('correlate', [True, False])
@pytest.mark.parametrize('adapt_to_errors', [False, True])
def test_thirteen_components_equal_narrow_runs(correlate, adapt_to_errors):
    import vegas_b200 as vegas
    from vegas_b200._gv import gv

    def run(cols, nitn=3):
        integ = vegas.Integrator(3 * [[0., 1.]], neval=20000, seed=11, correlate_integrals=correlate,
                                 adapt_to_errors=adapt_to_errors)
        f = vegas.lbatchintegrand(lambda x: _components(x)[:, cols])
        r = integ(f, nitn=nitn)
        last = r.itn_results[-1]
        return np.asarray(gv.mean(last), float), np.asarray(gv.evalcov(last), float), np.array(integ.sigf), np.array(integ.map.grid)

    full = list(range(13))
    mean, cov, sigf, grid = run(full)
    for cols in ([0, 1, 2, 3, 4, 5, 6, 7], [0, 8, 9, 10, 11, 12], [0, 5, 9, 12]):
        m2, c2, s2, g2 = run(cols)
        np.testing.assert_allclose(mean[cols], m2, rtol=1e-12, atol=1e-15)
        if correlate:
            np.testing.assert_allclose(cov[np.ix_(cols, cols)], c2, rtol=1e-9, atol=1e-18)
        else:
            np.testing.assert_allclose(np.diag(cov)[cols], np.diag(c2), rtol=1e-9, atol=1e-18)
        # component 0 drives the adaptation in both: the same stratification and map after 3 iterations (the
        # kernel instantiations differ in summation order, and a cube's variance is a difference of sums)
        np.testing.assert_allclose(sigf, s2, rtol=1e-9)
        np.testing.assert_allclose(grid, g2, rtol=1e-10)
    exact = np.array([0.5, 1 / 3., 0.25, np.sin(1.), 1 - np.exp(-1.), 0.25, 0.25, (1 - np.cos(3.)) / 3., 1 / 3., 1., 0.25, 1., 0.125])
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(mean - exact) < 5 * sd + 1e-12)
