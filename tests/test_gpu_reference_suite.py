"""The reference's own ``Integrator`` tests (``/root/reference/tests/test_vegas.py``, class
``TestIntegrator``; line numbers cited per test), run against ``vegas_b200`` through the public
API only: same integrands, same settings, same assertions.  They exercise the drop-in boundary --
plain Python / lbatch / rbatch integrands evaluated on host copies of GPU-generated samples, array-
and dictionary-valued results, ``uses_jac``, ``adapt_to_errors``, ``beta=0``, ``sample()``,
``random()``, tolerances, ``restratify``.  (Statistical assertions: the Philox stream differs from
the reference's PCG64, so seeds are fixed here and the bounds are the reference's.)"""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

REGION = [[0, math.pi], [-math.pi / 2., math.pi / 2.]]


def _v():
    import torch
    assert torch.cuda.is_available()
    import vegas_b200
    return vegas_b200


def _sincos(x):
    return (math.sin(x[0]) ** 2 + math.cos(x[1]) ** 2) / math.pi ** 2


def test_scalar_and_batch():
    """tests:941-983: scalar function and batch class integrate sin^2 + cos^2 to 1"""
    vegas = _v()
    r = vegas.Integrator(REGION, seed=1)(_sincos, neval=10000)
    assert abs(r.mean - 1.) < 5 * r.sdev and r.Q > 1e-3 and r.sdev < 1e-3

    @vegas.batchintegrand
    class f_batch:
        def __call__(self, x):
            return (np.sin(x[:, 0]) ** 2 + np.cos(x[:, 1]) ** 2) / math.pi ** 2
    r = vegas.Integrator(REGION, seed=2)(f_batch(), neval=10000)
    assert abs(r.mean - 1.) < 5 * r.sdev and r.Q > 1e-3 and r.sdev < 1e-3


def test_minimize_mem_is_accepted():
    """tests:985-1010: minimize_mem=True gives the same summary as False (here: the flag is a no-op,
    sigf always lives in HBM)"""
    vegas = _v()

    @vegas.batchintegrand
    def f(x):
        return (np.sin(x[:, 0]) ** 2 + np.cos(x[:, 1]) ** 2) / math.pi ** 2
    r = vegas.Integrator(REGION, minimize_mem=True, seed=3)(f, neval=10000)
    assert abs(r.mean - 1.) < 5 * r.sdev and r.Q > 1e-3 and r.sdev < 1e-3
    r2 = vegas.Integrator(REGION, minimize_mem=False, seed=3)(f, neval=10000)
    assert [str(a) for a in r2.itn_results] == [str(a) for a in r.itn_results]


def test_exceptions_propagate():
    """tests:1012-1029: exceptions raised inside the integrand come out unchanged"""
    vegas = _v()

    def f(x):
        return _sincos(x) / 0.0
    with pytest.raises(ZeroDivisionError):
        vegas.Integrator(REGION)(f, neval=100)

    @vegas.batchintegrand
    def fb(x):
        d = 1 / 0.
        return (np.sin(x[:, 0]) ** 2 + np.cos(x[:, 1]) ** 2) / d
    with pytest.raises(ZeroDivisionError):
        vegas.Integrator(REGION)(fb, neval=100)

    @vegas.batchintegrand
    def fnan(x):
        return np.full(x.shape[0], np.nan)
    with pytest.raises(ValueError):                      # pyx:2133-2134
        vegas.Integrator(REGION)(fnan, neval=100)


def test_beta0_and_adapt_to_errors():
    """tests:1031-1068: beta=0, adapt_to_errors, and both"""
    vegas = _v()

    @vegas.batchintegrand
    def fb(x):
        return (np.sin(x[:, 0]) ** 2 + np.cos(x[:, 1]) ** 2) / math.pi ** 2
    r = vegas.Integrator(REGION, beta=0.0, seed=4)(fb, neval=10000)
    assert abs(r.mean - 1.) < 5 * r.sdev and r.Q > 0.5e-3 and r.sdev < 1e-3
    for kw in (dict(adapt_to_errors=True), dict(adapt_to_errors=True, beta=0.0)):
        r = vegas.Integrator(REGION, seed=5, **kw)(_sincos, neval=10000)
        assert abs(r.mean - 1.) < 5 * r.sdev and r.Q > 1e-3 and r.sdev < 1e-3


def test_random_batch_and_random():
    """tests:1070-1100: re-summing integ.random_batch() / integ.random() reproduces the integral"""
    vegas = _v()

    def f(x):
        return math.exp(-100. * sum((x[d] - 0.5) ** 2 for d in range(4)))

    def fv(x):
        return np.exp(-100. * np.sum((x - 0.5) ** 2, axis=1))
    integ = vegas.Integrator(4 * [[0, 1]], seed=6)
    integ(f, nitn=10, neval=1000)
    result = integ(f, nitn=1, neval=1000, adapt=False)
    integral = sum(wgt.dot(fv(x)) for x, wgt in integ.random_batch())
    assert abs(result.mean - integral) < 5 * result.sdev

    def g(x):
        return x[0] ** 2 + x[1] ** 3
    integ = vegas.Integrator(2 * [[0, 2]], seed=7)
    integ(g, nitn=10, neval=100)
    result = integ(g, nitn=1, neval=100, adapt=False)
    integral = sum(wgt * g(x) for x, wgt in integ.random())
    assert abs(result.mean - integral) < 5 * result.sdev


def test_sample():
    """tests:1102-1152: integ.sample() for dictionary and array regions, both batch modes"""
    vegas = _v()
    neval, nitn, exact = 1000, 5, 1 * (8 - 1) * (81 - 16)

    @vegas.rbatchintegrand
    def f(x):
        return 2 * x['s'] * 3 * x['v'][0, 0] ** 2 * 4 * x['v'][1, 0] ** 3
    itg = vegas.Integrator(dict(s=(0., 1.), v=[[(1., 2.)], [(2., 3.)]]), neval=neval, nitn=nitn, seed=8)
    rv = itg(f)
    w, x = itg.sample(nbatch=nitn * itg.last_neval, mode='rbatch')
    assert len(w) == nitn * itg.last_neval
    assert abs(np.sum(w * f(x)) - exact) < 5 * rv.sdev

    @vegas.lbatchintegrand
    def f(x):
        return 2 * x['s'] * 3 * x['v'][:, 0, 0] ** 2 * 4 * x['v'][:, 1, 0] ** 3
    itg = vegas.Integrator(dict(s=(0., 1.), v=[[(1., 2.)], [(2., 3.)]]), neval=neval, nitn=nitn, seed=9)
    rv = itg(f)
    w, x = itg.sample(nbatch=nitn * itg.last_neval, mode='lbatch')
    assert len(w) == nitn * itg.last_neval
    assert abs(np.sum(w * f(x)) - exact) < 5 * rv.sdev

    @vegas.rbatchintegrand
    def f(x):
        return 2 * x[0, 0] * 3 * x[0, 1] ** 2 * 4 * x[0, 2] ** 3
    itg = vegas.Integrator([[(0., 1.), (1., 2.), (2., 3.)]], neval=neval, nitn=nitn, seed=10)
    rv = itg(f)
    w, x = itg.sample(nbatch=nitn * itg.last_neval, mode='rbatch')
    assert abs(np.sum(w * f(x)) - exact) < 5 * rv.sdev

    @vegas.lbatchintegrand
    def f(x):
        return 2 * x[:, 0, 0] * 3 * x[:, 0, 1] ** 2 * 4 * x[:, 0, 2] ** 3
    itg = vegas.Integrator([[(0., 1.), (1., 2.), (2., 3.)]], neval=neval, nitn=nitn, seed=11)
    rv = itg(f)
    w, x = itg.sample(nbatch=nitn * itg.last_neval, mode='lbatch')
    assert abs(np.sum(w * f(x)) - exact) < 5 * rv.sdev


def test_multi_integrands():
    """tests:1155-1196: array-valued integrands (batch class and scalar), correlated sums"""
    vegas = _v()

    def f_s(x):
        return math.exp(-100. * sum((x[d] - 0.5) ** 2 for d in range(4)))

    def f_multi_s(x):
        f = f_s(x)
        return [[f, f * x[0]]]

    @vegas.batchintegrand
    class f_multi_v:
        def __call__(self, x):
            x = np.asarray(x)
            f = np.empty((x.shape[0], 1, 2), float)
            f[:, 0, 0] = np.exp(-100. * np.sum((x - 0.5) ** 2, axis=1))
            f[:, 0, 1] = x[:, 0] * f[:, 0, 0]
            return f
    I = vegas.Integrator(4 * [[0, 1]], seed=12)
    I(f_s, neval=1000, nitn=10)
    for r in [I(f_multi_v(), nitn=10), I(f_multi_s, nitn=10)]:
        ratio = r[0, 1] / r[0, 0]
        assert abs(ratio.mean - 0.5) < 5 * ratio.sdev and ratio.sdev < 1e-2

    def f(x):
        f1 = np.sin(x[0]) * x[1]
        f2 = np.cos(x[1]) * x[0]
        return [f1 + f2, f1, f2]
    integ = vegas.Integrator([(0, 1), (0, 1)], seed=13)
    integ(f, neval=1e3, nitn=5)
    rs, r1, r2 = integ(f, neval=1e3, nitn=5)
    diff, r12 = rs - r1 - r2, r1 + r2
    assert diff.mean / r12.mean < 1e-7 and diff.sdev / r12.mean < 1e-7


def test_adaptive():
    """tests:1198-1211: adaptation reduces the error of the sharp 4-D Gaussian by more than 30x"""
    vegas = _v()

    def f(x):
        return math.exp(-100. * sum((x[i] - 0.5) ** 2 for i in range(4))) * 1013.2118364296088
    I = vegas.Integrator(4 * [[0, 1]], seed=14)
    r0 = I(f, neval=10000, nitn=10)
    r1 = I(f, neval=10000, nitn=10)
    assert r0.itn_results[0].sdev / 30 > r1.itn_results[-1].sdev
    assert r0.itn_results[0].sdev < 1. and r1.itn_results[-1].sdev < 0.01 and r1.Q > 1e-3


def test_dictintegrand():
    """tests:1213-1250: dictionary-valued integrands, scalar and batch"""
    vegas = _v()

    def f(x):
        return dict(a=x[0] + x[1], b=[[x[0] ** 2 * 3., x[1] ** 3 * 4.]])

    @vegas.batchintegrand
    def fb(x):
        ans = dict(a=x[:, 0] + x[:, 1], b=np.empty((x.shape[0], 1, 2), float))
        ans['b'][:, 0, 0] = x[:, 0] ** 2 * 3.
        ans['b'][:, 0, 1] = x[:, 1] ** 3 * 4.
        return ans
    for fcn, seed in ((f, 15), (fb, 16)):
        r = vegas.Integrator(2 * [[0, 1]], seed=seed)(fcn, neval=1000)
        for v in (r['a'], r['b'][0, 0], r['b'][0, 1]):
            assert abs(v.mean - 1.) < 5. * v.sdev and v.sdev < 1e-2
        assert r.dof == 27


def test_tol():
    """tests:1252-1270: rtol / atol stop the iterations early"""
    vegas = _v()

    def f(x):
        return 10 * np.exp(-100. * x[0]) * 100.
    for args, nitn in [(dict(), 2), (dict(rtol=0.5), 1), (dict(rtol=0.0001), 2), (dict(atol=0.5 * 10), 1),
                       (dict(atol=0.0001 * 10), 2)]:
        I = vegas.Integrator([[0, 1.]], neval=1000, nitn=2, seed=17, **args)
        assert I(f).nitn == nitn


def test_uses_jac():
    """tests:1298-1345: with uses_jac=True, f = 1/prod(jac) integrates to exactly 1 for every
    integrand type and output structure"""
    vegas = _v()
    integ = vegas.Integrator(2 * [[0, 2.]], seed=18)

    def check(f, mode):
        r = integ(f, nitn=1, neval=10, uses_jac=True)
        ans = r[0, 0] if mode == 'array' else (r['a'] if mode == 'dict' else r)
        assert abs(ans.mean - 1.) < 1e-7
    check(lambda x, jac: 1. / np.prod(jac), 'scalar')
    check(lambda x, jac: [[1. / np.prod(jac)]], 'array')
    check(lambda x, jac: dict(a=1. / np.prod(jac)), 'dict')
    check(vegas.rbatchintegrand(lambda x, jac: 1. / np.prod(jac, axis=0)), 'scalar')
    check(vegas.rbatchintegrand(lambda x, jac: [[1. / np.prod(jac, axis=0)]]), 'array')
    check(vegas.rbatchintegrand(lambda x, jac: dict(a=1. / np.prod(jac, axis=0))), 'dict')
    check(vegas.batchintegrand(lambda x, jac: 1. / np.prod(jac, axis=-1)), 'scalar')
    check(vegas.batchintegrand(lambda x, jac: [[1. / np.prod(jac, axis=-1)]]), 'array')
    check(vegas.batchintegrand(lambda x, jac: dict(a=1. / np.prod(jac, axis=-1))), 'dict')


def test_correlate():
    """tests:1347-1355: correlate_integrals switches the covariances on and off"""
    vegas = _v()
    from vegas_b200._gv import gv

    def f(x):
        return [np.prod(x), np.prod(x) ** 2]
    a, b = vegas.Integrator(2 * [[0, 1]], correlate_integrals=True, seed=19)(f)
    assert gv.corr(a, b) > 0.0
    aa, bb = vegas.Integrator(2 * [[0, 1]], correlate_integrals=False, seed=20)(f)
    assert gv.corr(aa, bb) == 0.0


def test_restratify_integrator():
    """tests:1357-1380: restratify(integ, f) on three Gaussians whose structure is on axes 0 and 1"""
    vegas = _v()
    norm = np.sqrt((100 / np.pi) ** 3) / 3
    x0list = np.array([[0.23, 0.23, 0.45], [0.39, 0.39, 0.45], [0.74, 0.74, 0.45]])

    @vegas.lbatchintegrand
    def f(x):
        ans = 0
        for x0 in x0list:
            ans = ans + np.exp(-100 * np.sum((x[:, :] - x0[None, :]) ** 2, axis=1))
        return ans * norm
    integ = vegas.Integrator(3 * [[0, 1]], alpha=0.5, seed=123)
    nitn = 2
    integ(f, nitn=15, neval=4e3)
    r = integ(f, alpha=0, adapt=True, nitn=nitn)
    integ2 = vegas.restratify(integ, f)
    r2 = integ2(f, alpha=0, adapt=True, nitn=nitn)
    dr = r2 - r
    assert 5 * dr.sdev > dr.mean
    assert integ2.nstrat[0] > 5 * integ2.nstrat[-1] and integ2.nstrat[1] > 5 * integ2.nstrat[-1]
    assert integ.neval == integ2.neval


def test_save_extend(tmp_path):
    """tests:479-530: the save / saveall keywords pickle the running result (and the integrator);
    the unpickled pair continues the integration and extends the saved result"""
    import pickle
    vegas = _v()
    from vegas_b200._gv import gv

    @vegas.rbatchintegrand
    def g(p):
        return p[0] ** 2 * 1.5 / 8

    @vegas.rbatchintegrand
    def ga(p):
        return [p[0] ** 2 * 1.5, 1 + p[0] ** 2 * 1.5]

    @vegas.rbatchintegrand
    def gd(p):
        return dict(x2=p[0] ** 2 * 1.5, one=[[1 + p[0] ** 2 * 1.5]])
    fn = str(tmp_path / 'test-save.pkl')
    itg = vegas.Integrator(2 * [[-1, 1]], nitn=2, neval=100, seed=21)
    for _g in [g, ga, gd]:
        r = itg(_g, save=fn)
        with open(fn, 'rb') as ifile:
            r1 = pickle.load(ifile)
        assert str(r1) == str(r) and r1.summary() == r.summary()
        if _g is not g:
            assert str(gv.evalcorr(r1.flat[:])) == str(gv.evalcorr(r.flat[:]))
    for _g in [g, ga, gd]:
        r = itg(_g, saveall=fn)
        with open(fn, 'rb') as ifile:
            r1, itg1 = pickle.load(ifile)
        assert str(r1) == str(r) and r1.summary() == r.summary()
        assert itg1.settings() == itg.settings()
        np.testing.assert_allclose(list(itg1.sigf), list(itg.sigf))
        new_r = itg1(_g)
        r1.extend(new_r)
        assert r.nitn + new_r.nitn == r1.nitn
        assert r.sum_neval + new_r.sum_neval == r1.sum_neval


def test_ravg_pickle():
    """tests:438-477: results rebuilt with vegas.ravg (with every form of rescale), pickled, and passed
    through gvar's dumps/loads print the same numbers, summaries and correlations"""
    import pickle
    vegas = _v()
    from vegas_b200._gv import gv

    @vegas.rbatchintegrand
    def g(p):
        return p[0] ** 2 * 1.5 / 8

    @vegas.rbatchintegrand
    def ga(p):
        return [p[0] ** 2 * 1.5, 1 + p[0] ** 2 * 1.5]

    @vegas.rbatchintegrand
    def gd(p):
        return dict(x2=p[0] ** 2 * 1.5, one=[[1 + p[0] ** 2 * 1.5]])
    itg = vegas.Integrator(2 * [[-1, 1]], nitn=2, neval=100, seed=22)
    for _g in [g, ga, gd]:
        r = itg(_g)

        def same(rx, r=r):
            assert str(rx) == str(r) and rx.summary() == r.summary()
            assert abs(rx.chi2 - r.chi2) < 1e-7 * max(1., abs(r.chi2))
            if _g is not g:
                assert str(gv.evalcorr(rx.flat[:])) == str(gv.evalcorr(r.flat[:]))
        same(vegas.ravg(r.itn_results))
        same(vegas.ravg(r.itn_results, rescale=r.itn_results[0]))
        same(vegas.ravg(r.itn_results, rescale=gv.mean(r.itn_results[0])))
        same(vegas.ravg(r))
        d3 = pickle.dumps(r)
        same(r)                                   # r unchanged by pickling
        same(pickle.loads(d3))
        d4 = gv.dumps(r)
        same(r)
        same(gv.loads(d4))
        r5 = vegas.ravg(r, weighted=False)
        assert str(r5) != str(r)
        same(r5, gv.loads(gv.dumps(r5)))


def test_volume():
    """tests:774-789: constants integrate to the volume exactly (zero variance), scalar and array"""
    vegas = _v()
    r = vegas.Integrator([[-1, 1], [0, 4]], seed=23)(lambda x: 2.)
    np.testing.assert_allclose(r.mean, 16, rtol=1e-6)
    assert r.sdev < 1e-6
    r = vegas.Integrator([[-1, 1], [0, 4]], seed=24)(lambda x: [-1., 2.])
    np.testing.assert_allclose(r[0].mean, -8, rtol=5e-2)
    np.testing.assert_allclose(r[1].mean, 16, rtol=5e-2)
    assert r[0].sdev < 1e-6 and r[1].sdev < 1e-6
