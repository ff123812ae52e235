"""Pin the CPU oracle (oracle/vegas_oracle.c + oracle/oracle.py) against the REFERENCE:

* golden fixtures under tests/golden/ (produced by tests/golden/make_golden.py from the unmodified
  reference module fed a recorded uniform stream), and
* the reference's own known-answer tests (tests/test_vegas.py line numbers cited per test).

CPU only.  The GPU parity tests then compare the CUDA engine with this oracle."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden.cases import CASES, integrand

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 2e-13


def _load(name):
    return np.load(os.path.join(HERE, 'golden', 'ref_%s.npz' % name))


@pytest.mark.parametrize('name', sorted(CASES))
def test_iterations_match_reference(name):
    spec, G = CASES[name], _load(name)
    v = O.Vegas(spec['limits'], **spec['kw'])
    assert list(v.nstrat) == list(G['nstrat']) and list(v.map.ninc) == list(G['ninc'])
    assert v.nhcube == int(G['nhcube']) and v.min_neval_hcube == int(G['min_neval_hcube'])
    rng = np.random.default_rng(spec['seed'])
    rows = []

    def uniforms(h0, nh):
        rows.append(int(nh.sum()))
        return rng.random((int(nh.sum()), v.dim))

    f = integrand(spec['f'])
    for i in range(spec['nitn']):
        mean, var = v.iterate(f, uniforms)
        assert v.last_neval == int(G['itn%d_last_neval' % i])
        np.testing.assert_allclose(mean, G['itn%d_mean' % i], rtol=TOL, atol=1e-300)
        cov = G['itn%d_cov' % i]
        if np.ndim(var) == 2:
            np.testing.assert_allclose(var, cov, rtol=1e-11, atol=1e-18 * np.abs(cov).max())
        else:
            np.testing.assert_allclose(var, np.diag(cov), rtol=1e-11)
        if len(v.sigf):
            np.testing.assert_allclose(v.sigf, G['itn%d_sigf' % i], rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(v.sum_sigf, float(G['itn%d_sum_sigf' % i]), rtol=TOL)
            if spec['kw'].get('beta', 0.75) > 0 and not spec['kw'].get('adapt_to_errors', False):
                assert tuple(v.neval_hcube_range) == tuple(G['itn%d_range' % i])
        v.adapt_map()
        g = G['itn%d_grid' % i]
        for d in range(v.dim):
            n = v.map.ninc[d] + 1
            np.testing.assert_allclose(v.map.grid[d, :n], g[d, :n], rtol=1e-12, atol=1e-15)
    # the oracle cuts batches exactly where the reference did (one generator call per batch)
    assert rows == list(G['batch_rows'])


def test_map_invmap_jac1d_bit_exact():
    G = _load('map')
    m = O.Map([G['grid0'], G['grid1'], G['grid2']])
    x, jac = m.map(G['y'])
    assert np.array_equal(x, G['x']) and np.array_equal(jac, G['jac'])
    assert np.array_equal(m.jac1d(G['y']), G['jac1d'])
    mi = O.Map([G['gridi0'], G['gridi1'], G['gridi2']])
    y, jac2 = mi.invmap(G['inv_x'])
    assert np.array_equal(y, G['inv_y']) and np.array_equal(jac2, G['inv_jac'])


def test_training_and_adapt_match_reference():
    G = _load('map')
    m = O.Map([[0, 2], [-1, 1]], ninc=[50, 33])
    for i, alpha in enumerate((1.5, 0.5, -1.0)):
        m.add_training_data(G['train%d_y' % i], G['train%d_f' % i])
        for d in range(2):
            n = m.ninc[d]
            assert np.array_equal(m.sum_f[d, :n], G['train%d_sum_f' % i][d, :n])
            assert np.array_equal(m.n_f[d, :n], G['train%d_n_f' % i][d, :n])
        m.adapt(alpha=alpha)
        g = G['train%d_grid' % i]
        for d in range(2):
            n = m.ninc[d] + 1
            np.testing.assert_allclose(m.grid[d, :n], g[d, :n], rtol=1e-14, atol=1e-16)
    m.adapt(ninc=[20, 7])
    for d, n in enumerate((21, 8)):
        np.testing.assert_allclose(m.grid[d, :n], G['regrid'][d, :n], rtol=1e-14, atol=1e-16)


# ----------------------------------------------------------------------------- reference known answers
def test_ref_map_known_answers():
    """reference tests/test_vegas.py:76-112"""
    m = O.Map([[0, 1, 3], [-2, 0, 6]])
    y = np.array([[0, 0], [0.25, 0.25], [0.5, 0.5], [0.75, 0.75], [1.0, 1.0]])
    x, jac = m.map(y)
    np.testing.assert_allclose(x, [[0, -2], [0.5, -1], [1, 0], [2, 3], [3, 6]])
    np.testing.assert_allclose(jac, [8, 8, 48, 48, 48])
    yi, ji = m.invmap(x)
    np.testing.assert_allclose(yi, y)
    np.testing.assert_allclose(ji, jac)


def test_ref_init_regrid():
    """reference tests/test_vegas.py:39-63"""
    m = O.Map([[0, 1], [-2, 4]], ninc=2)
    np.testing.assert_allclose(m.grid, [[0, 0.5, 1.], [-2., 1., 4.]])
    m = O.Map([[0, 0.4, 1], [-2, 0., 4]], ninc=4)
    np.testing.assert_allclose(m.grid, [[0, 0.2, 0.4, 0.7, 1.], [-2., -1., 0., 2., 4.]])
    np.testing.assert_allclose(m.inc, [[0.2, 0.2, 0.3, 0.3], [1, 1, 2, 2]])


def test_ref_training_adapt_analytic():
    """reference tests/test_vegas.py:142-207: adapt converges to the analytic optimal grids"""
    g = 1. / 3. ** 0.5
    ygauss = [(1 - g) / 4., (1 + g) / 4, (3 - g) / 4, (3 + g) / 4.]
    m = O.Map([[0, 2]], ninc=2)
    y = np.array([[yi] for yi in ygauss])
    for _ in range(60):
        x, jac = m.map(y)
        m.add_training_data(y, x[:, 0] ** 2 * jac)
        m.adapt(alpha=2.)
    np.testing.assert_allclose(m.grid, [[0, 2. / 2. ** (1. / 3.), 2.]])
    for alpha, nit in ((2., 60), (-2., 20)):
        m = O.Map([[0, 2], [0, 4]], ninc=2)
        y = np.array([[yi, yj] for yi in ygauss for yj in ygauss])
        for _ in range(nit):
            x, jac = m.map(y)
            m.add_training_data(y, x[:, 0] * x[:, 1] ** 2 * jac)
            m.adapt(alpha=alpha)
        np.testing.assert_allclose(m.grid, [[0, 2. * 2. ** (-0.5), 2.], [0, 4. * 2 ** (-1. / 3.), 4.]])


def test_ref_strata_integers():
    """reference tests/test_vegas.py:561-595, 733-769: nstrat / ninc / neval / min_neval_hcube"""
    s = O.strata(234, 2, None)
    assert list(s['nstrat']) == [5, 5] and list(s['ninc']) == [20, 20]
    s = O.strata(1000, 2, None, nstrat=[1, 1])
    assert list(s['ninc']) == [100, 100] and s['min_neval_hcube'] == 1000
    s = O.strata(None, 2, None, nstrat=[10, 11])
    assert list(s['ninc']) == [80, 88] and s['neval'] == 880 and s['min_neval_hcube'] == 2
    s = O.strata(2000, 2, None, nstrat=[7, 9])
    assert list(s['ninc']) == [196, 198] and s['min_neval_hcube'] == 7
    assert list(O.strata(3100, 2, None)['nstrat']) == [20, 19]
    assert list(O.strata(3500, 2, None)['nstrat']) == [21, 20]
    s = O.strata(1e4, 4, None)
    assert list(s['nstrat']) == [6, 6, 6, 5] and list(s['ninc']) == [996, 996, 996, 1000]
    s = O.strata(1e8, 8, None)
    assert list(s['nstrat']) == [8, 8, 8, 8, 8, 7, 7, 7] and s['nhcube'] == 11239424


def test_ref_weighted_average_known_answer():
    """reference tests/test_vegas.py:216-236"""
    mean, sdev, chi2, dof, Q = O.wavg([1., 2., 3.], [1., 4., 9.])
    np.testing.assert_allclose([mean, sdev, chi2, Q], [1.346938775510204, 0.8571428571428571,
                                                       0.5306122448979592, 0.7669711269557102])
    assert dof == 2


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (philox.h kat_vectors)"""
    assert O.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert O.philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert O.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    u = O.philox_uniforms(12345, 3, 5, 7, np.array([3, 2], np.int64))
    assert u.shape == (5, 5) and u.min() >= 0 and u.max() < 1


def test_oracle_vs_compiled_reference_live():
    """where oracle/_ref exists (build container and GPU box): one more live cross-check with a fresh
    uniform stream"""
    import sys
    ref_dir = os.path.join(os.path.dirname(HERE), 'oracle', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'vegas')):
        pytest.skip('oracle/_ref not built')
    import subprocess
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import gvar, vegas
from oracle import oracle as O
gvar.ranseed(1)
rng = np.random.default_rng(99)
f = lambda x: np.exp(-50. * np.sum((x - 0.3) ** 2, axis=1))
I = vegas.Integrator(3 * [[0., 1.]], neval=5000, ran_array_generator=lambda s: rng.random(s))
r = I(vegas.lbatchintegrand(f), nitn=4)
rng2 = np.random.default_rng(99)
v = O.Vegas(3 * [[0., 1.]], neval=5000)
ms, vs = [], []
for i in range(4):
    m, var = v.iterate(f, lambda h0, nh: rng2.random((int(nh.sum()), 3)))
    v.adapt_map(); ms.append(m[0]); vs.append(var[0, 0])
ref = [(x.mean, x.sdev ** 2) for x in r.itn_results]
for (a, b), c, d in zip(ref, ms, vs):
    assert abs(a - c) <= 2e-13 * abs(a) and abs(b - d) <= 1e-10 * abs(b), (a, c, b, d)
for d in range(3):          # rows are padded to the widest axis; the padding is uninitialised in the reference
    n = int(v.map.ninc[d]) + 1
    assert np.allclose(np.array(I.map.grid)[d, :n], v.map.grid[d, :n], rtol=1e-11, atol=1e-14)
print('ok')
''' % (os.path.join(os.path.dirname(HERE), 'oracle', 'gvar_shim'), ref_dir, os.path.dirname(HERE))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert out.returncode == 0 and 'ok' in out.stdout, out.stderr[-2000:]


def test_restratify_matches_reference():
    """vegas.restratify (src/vegas/__init__.py:1313-1419): the oracle's restatement of the auxiliary
    I/dI integrand, the weights and the new stratification vs the unmodified reference run recorded
    by tests/golden/make_golden_restratify.py (same injected uniforms)"""
    from tests.golden.cases import RESTRATIFY
    G = np.load(os.path.join(HERE, 'golden', 'ref_restratify.npz'))
    for name, spec in RESTRATIFY.items():
        v = O.Vegas(spec['limits'], alpha=0.0, correlate_integrals=False, **spec['kw'])
        assert list(v.nstrat) == list(G[name + '_old_nstrat'])
        v.map.grid = G[name + '_grid'][:, :v.map.grid.shape[1]].copy()
        v.sigf = G[name + '_sigf'].copy()
        v.sum_sigf = float(G[name + '_sum_sigf'])
        rng = np.random.default_rng(spec['seed'] + 1000)
        ndy = spec['ndy']
        fcn = O.profile_integrand(v.map, integrand(spec['f']), ndy)
        means, variances = [], []
        for i in range(spec['nitn']):
            mean, var = v.iterate(fcn, lambda h0, nh: rng.random((int(nh.sum()), v.dim)))
            means.append(mean)
            variances.append(np.asarray(var).reshape(-1))
        m, s2 = np.array(means), np.array(variances)
        w = 1. / s2
        avg = (m * w).sum(axis=0) / w.sum(axis=0)                 # independent components: plain weighted averages
        avar = 1. / w.sum(axis=0)
        np.testing.assert_allclose([avg[0], avar[0]], G[name + '_I'], rtol=1e-11)
        np.testing.assert_allclose(avg[1:].reshape(v.dim, ndy), G[name + '_dI_mean'], rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(avar[1:].reshape(v.dim, ndy), G[name + '_dI_var'], rtol=1e-9, atol=1e-300)
        weight = O.restratify_weights(avg[0], avg[1:].reshape(v.dim, ndy), ndy)
        np.testing.assert_allclose(weight, G[name + '_weight'], rtol=1e-8)
        new = O.restratify_nstrat(v.nstrat, G[name + '_weight'], **spec['opt'])
        assert list(new) == list(G[name + '_new_nstrat'])
