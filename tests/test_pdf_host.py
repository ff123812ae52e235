"""Host side of ``PDFIntegrator`` (no GPU): the built-in gvar pieces it needs (string GVars, ``PDF``,
``PDFStatistics``), the numpy twin of the integrand against what the unmodified reference computed
(``tests/golden/ref_pdf.npz``, generator ``tests/golden/make_golden_pdf.py``), and the result classes."""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vegas_b200 as vegas                            # noqa: E402
from vegas_b200 import _gvbuiltin as gvb              # noqa: E402
from vegas_b200._gv import gv                         # noqa: E402
from vegas_b200._integrand import VegasIntegrand      # noqa: E402
from tests.golden.cases import PDF_CASES, pdf_f       # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def test_string_gvars():
    for text, m, s in [('1.0(5)', 1.0, 0.5), ('13.5(8.6)', 13.5, 8.6), ('1(1)', 1., 1.), ('2.5(2.7)', 2.5, 2.7),
                       ('-1.35(86)e-2', -0.0135, 0.0086), ('12(3)', 12., 3.), ('1.2 +- 0.3', 1.2, 0.3), ('0.120(45)', 0.12, 0.045)]:
        g = gvb.gvar(text)
        assert abs(g.mean - m) < 1e-15 and abs(g.sdev - s) < 1e-15, text
    a = gvb.gvar([2 * ['1(1)']])
    assert a.shape == (1, 2) and a[0, 1].sdev == 1.
    d = gvb.gvar(dict(a='1(1)', b=[2 * ['2(2)']]))
    assert d['b'].shape == (1, 2) and d['a'].mean == 1. and d.size == 3
    assert str(gvb.gvar('1.0(5)')) == '1.00(50)'


def test_pdf_class():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(4, 4))
    cov = a @ a.T + 0.1 * np.eye(4)
    mean = rng.normal(size=4)
    pdf = gvb.PDF(gvb.gvar(mean, cov))
    np.testing.assert_allclose(pdf.vec_sig.T @ pdf.vec_sig, cov, rtol=1e-12)
    np.testing.assert_allclose(pdf.vec_isig.T @ pdf.vec_isig, np.linalg.inv(cov), rtol=1e-10)
    np.testing.assert_allclose(pdf.dp_dchiv, np.sqrt(np.linalg.det(cov)), rtol=1e-12)
    c = rng.normal(size=(7, 4))
    p = pdf.pflat(c, mode='lbatch')
    np.testing.assert_allclose(pdf.chiv(p, mode='lbatch'), c, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(pdf.pflat(c.T, mode='rbatch'), p.T, rtol=1e-14)
    from scipy.stats import multivariate_normal
    np.testing.assert_allclose(pdf.pdf(p, mode='lbatch'), multivariate_normal(mean, cov).pdf(p), rtol=1e-10)
    # layouts: scalar, array, dictionary
    assert gvb.PDF(gvb.gvar(1, 2)).shape == () and np.ndim(gvb.PDF(gvb.gvar(1, 2)).sample()) == 0
    pd = gvb.PDF(gvb.gvar(dict(a='1(1)', b=['2(2)', '3(1)'])))
    s = pd.sample()
    assert pd.size == 3 and pd.shape is None and s['b'].shape == (2,)
    lb = pd._unflatten(np.arange(6.).reshape(2, 3), mode='lbatch')
    assert lb['a'].shape == (2,) and lb['b'].shape == (2, 2) and lb['b'][1, 0] == 4.
    rb = pd._unflatten(np.arange(6.).reshape(3, 2), mode='rbatch')
    assert rb['b'].shape == (2, 2) and rb['b'][1, 0] == 4.
    # svdcut lifts a nearly singular correlation matrix
    sing = gvb.gvar([0., 0.], [[1., 1 - 1e-14], [1 - 1e-14, 1.]])
    assert gvb.PDF(sing, svdcut=1e-6).dp_dchiv > 1e3 * gvb.PDF(sing, svdcut=1e-15).dp_dchiv


def test_pdfstatistics():
    z = gvb.gvar
    # moments of N(1, 2^2): <x>=1, <x^2>=5, <x^3>=13, <x^4>=73
    st = gvb.PDFStatistics(moments=[z(1, 1e-3), z(5, 1e-3), z(13, 1e-3), z(73, 1e-3)])
    assert abs(st.mean.mean - 1) < 1e-12 and abs(st.sdev.mean - 2) < 1e-12
    assert abs(st.skew.mean) < 1e-12 and abs(st.ex_kurt.mean) < 1e-12 and st.skew.sdev > 0
    # histogram of the same distribution: median 1, +/- 2
    from scipy.stats import norm
    bins = 1 + np.linspace(-6, 6, 13)
    cdf = norm(1, 2).cdf(bins)
    count = np.concatenate([[cdf[0]], np.diff(cdf), [1 - cdf[-1]]])
    st = gvb.PDFStatistics(moments=[z(1, 1e-3), z(5, 1e-3)], histogram=(bins, gvb.gvar(count, 1e-6 * np.ones(14))))
    assert abs(gvb.mean(st.median.loc) - 1) < 1e-6
    assert abs(gvb.mean(st.median.plus) - 2) < 0.05 and abs(gvb.mean(st.median.minus) - 2) < 0.05
    assert 'median' in str(st) and 'mean' in str(st)


def test_numpy_twin_of_the_integrand_vs_reference():
    """PDFIntegrator._f_lbatch + the built-in PDF against the rows the reference's own _f_lbatch produced"""
    G = np.load(os.path.join(HERE, 'golden', 'ref_pdf.npz'))
    for name, spec in PDF_CASES.items():
        pdf = gv.PDF(gv.gvar(spec['mean'], spec['cov']), svdcut=1e-15)
        np.testing.assert_allclose(pdf.vec_sig, G[name + '_vec_sig'], rtol=1e-13, atol=1e-15)
        f = VegasIntegrand(vegas.lbatchintegrand(pdf_f), None, False, pdf._unflatten(pdf.meanflat), False)
        ans = vegas.PDFIntegrator._f_lbatch(G[name + '_theta'], f, pdf, None, spec['scale'], spec['adapt_to_pdf'])
        assert [str(k) for k in ans.keys()] == list(G[name + '_keys'])
        rows = np.concatenate([np.asarray(ans[k], float).reshape(len(G[name + '_theta']), -1) for k in ans], axis=1)
        np.testing.assert_allclose(rows, G[name + '_rows'], rtol=1e-13, atol=1e-300)


def _fake_results(keys_shapes, nitn=3, seed=0):
    rng = np.random.default_rng(seed)
    itn = []
    for _ in range(nitn):
        d = gv.BufferDict()
        for k, shape in keys_shapes:
            n = int(np.prod(shape, dtype=int))
            vals = gv.gvar(1 + 0.01 * rng.normal(size=n), 0.01 * np.ones(n))
            d[k] = vals[0] if shape == () else vals.reshape(shape)
        itn.append(d)
    return vegas.ravg(itn)


def test_result_classes():
    res = _fake_results([('pdf', ()), ('f(p)*pdf', ())])
    ev = vegas.PDFEV(res)
    assert isinstance(ev, gv.GVar) and abs(ev.mean - 1) < 0.05 and ev.sdev > 0
    assert ev.pdfnorm is res['pdf'] or ev.pdfnorm.mean == res['pdf'].mean
    assert ev.Q == res.Q and len(ev.itn_results) == 3 and 'itn' in ev.summary()
    ev2 = pickle.loads(pickle.dumps(ev))
    assert isinstance(ev2, vegas.PDFEV) and str(ev2) == str(ev)
    ev.extend(vegas.PDFEV(_fake_results([('pdf', ()), ('f(p)*pdf', ())], seed=1)))
    assert len(ev.results.itn_results) == 6
    res = _fake_results([('pdf', ()), ('f(p)*pdf', (2, 2))])
    eva = vegas.PDFEVArray(res)
    assert eva.shape == (2, 2) and isinstance(eva[0, 1], gv.GVar) and eva.pdfnorm.mean == res['pdf'].mean
    assert pickle.loads(pickle.dumps(eva)).shape == (2, 2)
    res = _fake_results([('pdf', ()), (('f(p)*pdf', 'a'), ()), (('f(p)*pdf', 'b'), (3,))])
    evd = vegas.PDFEVDict(res)
    assert list(evd.keys()) == ['a', 'b'] and evd['b'].shape == (3,) and abs(evd['a'].mean - 1) < 0.05
    assert isinstance(pickle.loads(pickle.dumps(evd)), vegas.PDFEVDict)
    assert isinstance(vegas.ravg(evd, weighted=False), vegas.PDFEVDict)
    assert vegas.PDFIntegrator._make_ans(res).__class__ is vegas.PDFEVDict
    try:
        evd.keys_
        raise AssertionError('attribute error expected')
    except AttributeError:
        pass
