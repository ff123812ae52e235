"""``PDFIntegrator`` on the GPU: the device kernels of the change of variables against the numpy twin
of the reference's wrapper (``src/vegas/__init__.py:593-640``), and the reference's own
``test_PDFIntegrator`` cases (``/root/reference/tests/test_vegas.py:1410-1870``; line numbers cited per
test) through the public API: same parameters, same functions, same assertions.  (Statistical
assertions: the Philox stream differs from the reference's generator, so the engine seed is fixed
here and the bounds are the reference's.)"""
import collections
import io
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _v():
    import torch
    assert torch.cuda.is_available()
    import vegas_b200
    from vegas_b200._gv import gv
    return vegas_b200, gv


G2 = ([1., 2.], [[1, .1], [.1, 4]])


def test_pdf_map_kernels_vs_numpy_twin():
    """k_pdf_map / k_pdf_weight against PDFIntegrator._f_lbatch (the reference's formulas) on the same theta:
    p, the weight and the assembled rows to 1e-13; both column orders; an array-valued f."""
    import torch
    vegas, gv = _v()
    from vegas_b200 import _lib
    from vegas_b200._integrand import VegasIntegrand
    rng = np.random.default_rng(5)
    a = rng.normal(size=(5, 5))
    cov = a @ a.T + 0.3 * np.eye(5)
    g = gv.gvar(rng.normal(size=5), cov)
    pdf = gv.PDF(g)
    rows, scale = 4097, 1.7
    theta = rng.uniform(-1.45, 1.45, size=(rows, 5))
    f = VegasIntegrand(vegas.lbatchintegrand(lambda p: np.stack([p[:, 0] * p[:, 1], np.cos(p[:, 2]), p[:, 4] ** 2], axis=1)),
                       None, False, pdf.sample(mode=None), False)
    ctx = _lib.Context(None)
    dev = ctx.device
    th = torch.from_numpy(theta).to(dev)
    mean = torch.from_numpy(pdf.meanflat.copy()).to(dev)
    vs = torch.from_numpy(np.ascontiguousarray(pdf.vec_sig)).to(dev)
    p = torch.empty_like(th)
    w = torch.empty(rows, dtype=torch.float64, device=dev)
    ctx.pdf_map(th, scale, pdf.dp_dchiv, True, mean, vs, p, w)
    chiv = scale * np.tan(theta)
    np.testing.assert_allclose(p.cpu().numpy(), pdf.pflat(chiv, mode='lbatch'), rtol=1e-13, atol=1e-13)
    for adapt_to_pdf in (True, False):
        ref = vegas.PDFIntegrator._f_lbatch(theta, f, pdf, None, scale, adapt_to_pdf)
        np.testing.assert_allclose(w.cpu().numpy(), ref['pdf'], rtol=1e-13, atol=1e-300)
        fp = torch.from_numpy(np.ascontiguousarray(f.eval(p.cpu().numpy()))).to(dev)
        out = torch.empty((rows, 4), dtype=torch.float64, device=dev)
        ctx.pdf_weight(fp, w, out, adapt_to_pdf)
        oh = out.cpu().numpy()
        cols = (oh[:, 0], oh[:, 1:]) if adapt_to_pdf else (oh[:, 3], oh[:, :3])
        np.testing.assert_allclose(cols[0], ref['pdf'], rtol=1e-13, atol=1e-300)
        # (f is evaluated at the device's p, which differs from numpy's in the last bit: absolute floor)
        np.testing.assert_allclose(cols[1], ref['f(p)*pdf'], rtol=1e-12, atol=1e-15)
        assert list(ref.keys()) == (['pdf', 'f(p)*pdf'] if adapt_to_pdf else ['f(p)*pdf', 'pdf'])
    # without the Gaussian: the Jacobian alone
    ctx.pdf_map(th, scale, pdf.dp_dchiv, False, mean, vs, p, w)
    np.testing.assert_allclose(w.cpu().numpy(), np.prod(scale * (np.tan(theta) ** 2 + 1.), axis=1) * pdf.dp_dchiv, rtol=1e-13)


def test_nobatch():
    """tests:1411-1450: scalar, array- and dictionary-valued f(p) without batching"""
    vegas, gv = _v()
    g = gv.gvar(*G2)
    gev = vegas.PDFIntegrator(g, adapt=False, seed=1)

    def f(p):
        return p[0] + p[1]
    r = gev(f, nitn=1)
    assert isinstance(r, vegas.PDFEV)
    assert abs(r.mean - sum(g).mean) < 5 * r.sdev
    assert abs(r.pdfnorm.mean - 1) < 5 * r.pdfnorm.sdev

    def f(p):
        ff = p[0] + p[1]
        ff2 = ff * ff
        return [[ff, ff, ff2], [ff2, ff, ff2 * ff]]
    r = gev(f, nitn=1)
    assert isinstance(r, vegas.PDFEVArray) and r.shape == (2, 3)
    assert abs(r[0, 0].mean - sum(g).mean) < 5 * r[0, 0].sdev
    var = r[1, 0] - r[0, 0] ** 2
    assert abs(var.mean - sum(g).var) < 5. * var.sdev
    assert str(r[0, 0]) == str(r[0, 1]) == str(r[1, 1])
    assert str(r[0, 2]) == str(r[1, 0])
    diff3 = r[1, 2] - 3 * r[1, 0] * r[0, 0] + 2 * r[0, 0] ** 3
    assert abs(diff3.mean) < 5. * diff3.sdev

    def f(p):
        ff = p[0] + p[1]
        ff2 = ff * ff
        ans = collections.OrderedDict()
        ans[0] = ff
        ans[1] = [[ff, ff, ff2], [ff2, ff, ff2 * ff]]
        return ans
    r = gev(f, nitn=1)
    assert isinstance(r, vegas.PDFEVDict)
    assert str(r[0]) == str(r[1][0, 0]) == str(r[1][0, 1]) == str(r[1][1, 1])
    assert str(r[1][0, 2]) == str(r[1][1, 0])
    assert r[1].shape == (2, 3)
    assert abs(r[0].mean - sum(g).mean) < 5 * r[0].sdev


def test_rbatch_lbatch():
    """tests:1488-1590: the same with rbatch and lbatch integrands"""
    vegas, gv = _v()
    g = gv.gvar(*G2)
    gev = vegas.PDFIntegrator(g, adapt=False, seed=1)

    @vegas.rbatchintegrand
    def f(p):
        return p[0] + p[1]
    r = gev(f, nitn=1)
    assert abs(r.mean - sum(g).mean) < 5 * r.sdev

    @vegas.rbatchintegrand
    def f(p):
        ff = p[0] + p[1]
        ff2 = ff * ff
        return [[ff, ff, ff2], [ff2, ff, ff2 * ff]]
    r = gev(f, nitn=1)
    assert r.shape == (2, 3)
    var = r[1, 0] - r[0, 0] ** 2
    assert abs(var.mean - sum(g).var) < 5. * var.sdev

    @vegas.lbatchintegrand
    def f(p):
        ff = p[:, 0] + p[:, 1]
        ff2 = ff * ff
        return dict(a=ff, b=np.moveaxis(np.array([[ff, ff, ff2], [ff2, ff, ff2 * ff]]), -1, 0))
    r = gev(f, nitn=1)
    assert r['b'].shape == (2, 3)
    assert str(r['a']) == str(r['b'][0, 0]) == str(r['b'][1, 1])
    assert abs(r['a'].mean - sum(g).mean) < 5 * r['a'].sdev


def test_scalar_array_dict_param():
    """tests:1592-1650: param a single GVar, an array, a dictionary"""
    vegas, gv = _v()
    g = gv.gvar(1, 2)
    gev = vegas.PDFIntegrator(g, alpha=0, beta=0, seed=2)
    r = gev(vegas.rbatchintegrand(lambda p: p), nitn=1)
    assert abs(r.mean - g.mean) < 5 * r.sdev
    r = gev(vegas.lbatchintegrand(lambda p: p), nitn=1)
    assert abs(r.mean - g.mean) < 5 * r.sdev
    g = gv.gvar(*G2)
    gev = vegas.PDFIntegrator(g, alpha=0, beta=0, seed=2)
    r = gev(vegas.rbatchintegrand(lambda p: np.sum(p, axis=0)), nitn=1)
    assert abs(r.mean - sum(g).mean) < 5 * r.sdev
    r = gev(vegas.lbatchintegrand(lambda p: np.sum(p, axis=1)), nitn=1)
    assert abs(r.mean - sum(g).mean) < 5 * r.sdev
    g = dict(a=gv.gvar(1, 1), b=gv.gvar(*G2))
    gev = vegas.PDFIntegrator(g, alpha=0, beta=0, seed=2)
    r = gev(lambda p: np.sum(p['b'], axis=0), nitn=1)
    assert abs(r.mean - sum(g['b']).mean) < 5 * r.sdev
    r = gev(vegas.rbatchintegrand(lambda p: np.sum(p['b'], axis=0)), nitn=1)
    assert abs(r.mean - sum(g['b']).mean) < 5 * r.sdev
    r = gev(vegas.lbatchintegrand(lambda p: np.sum(p['b'], axis=1)), nitn=1)
    assert abs(r.mean - sum(g['b']).mean) < 5 * r.sdev


def test_change_pdf_and_adapt_to_pdf():
    """tests:1652-1686: a user PDF (shifted peak), and adaptation to f(p) pdf(p)"""
    vegas, gv = _v()
    g = gv.gvar(1, 2)
    gev = vegas.PDFIntegrator(g, alpha=0, beta=0, seed=3)

    def pdf(p):
        return np.exp(-(p - g.mean - 0.5 * g.sdev) ** 2 / 8) / np.sqrt(2 * np.pi * g.var)
    r = gev(vegas.rbatchintegrand(lambda p: [p, p ** 2]), pdf=pdf, nitn=2, adapt=True)
    assert abs(r[0].mean - g.mean - 0.5 * g.sdev) < 10 * r[0].sdev
    gev = vegas.PDFIntegrator(g, alpha=0, beta=0, seed=3)
    r = gev(vegas.lbatchintegrand(lambda p: np.moveaxis(np.array([p, p ** 2]), -1, 0)), pdf=pdf, nitn=2, adapt=True)
    assert abs(r[0].mean - g.mean - 0.5 * g.sdev) < 10 * r[0].sdev
    g2 = gv.gvar(*G2)
    gev = vegas.PDFIntegrator(g2, alpha=0, beta=0, adapt_to_pdf=False, seed=3)
    r = gev(vegas.rbatchintegrand(lambda p: p[0] + p[1]), nitn=2)
    assert abs(r.mean - sum(g2).mean) < 5 * r.sdev
    assert list(r.results.keys()) == ['f(p)*pdf', 'pdf']


def test_limit_scale():
    """tests:1688-1702: the probability inside 1 and 2 standard deviations, for two scales"""
    vegas, gv = _v()
    g = gv.gvar(1, 0.1)
    for scale in [1., 2.]:
        norm = vegas.PDFIntegrator(g, limit=1., scale=scale, seed=4)(neval=1000, nitn=5).pdfnorm
        assert abs(norm.mean - 0.682689492137) < 5 * norm.sdev
        norm = vegas.PDFIntegrator(g, limit=2., scale=scale, seed=4)(neval=1000, nitn=5).pdfnorm
        assert abs(norm.mean - 0.954499736104) < 5 * norm.sdev


def test_no_f_and_stats():
    """tests:1704-1712, 1837-1862: the norm alone; stats() reproduces the parameters' means and widths"""
    vegas, gv = _v()
    for g in [gv.gvar(1, 1), gv.gvar([2 * ['1(1)']]), gv.gvar(dict(a='1(1)', b=[2 * ['2(2)']]))]:
        gev = vegas.PDFIntegrator(g, seed=5)
        norm = gev(nitn=1).pdfnorm
        assert abs(norm.mean - 1) < 5 * norm.sdev
        st = gev.stats()
        gf = [g] if isinstance(g, gv.GVar) else (g.buf if hasattr(g, 'keys') else g.flat[:])
        sf = [st] if isinstance(st, gv.GVar) else (st.buf if hasattr(st, 'keys') else np.asarray(st).flat[:])
        for a, b in zip(gf, sf):
            assert abs(a.mean - b.mean) < 0.05 and abs(a.sdev - b.sdev) < 0.05 * a.sdev
    for gs in ['1(2)', [['1(2)'], ['2(3)']], dict(a='1(2)', b=['2(3)'])]:
        g = gv.gvar(gs)
        gev = vegas.PDFIntegrator(g, neval=4000, seed=6)
        gev()
        r = gev.stats(vegas.rbatchintegrand(lambda p: p), moments=True, histograms=True)
        if isinstance(g, gv.GVar):
            assert isinstance(r, gv.GVar)
            gf, rf, sf = [g], [r], [r.stats]
        elif hasattr(g, 'keys'):
            gf, rf, sf = g.buf, r.buf, r.stats.buf
        else:
            gf, rf, sf = g.flat[:], np.asarray(r).flat[:], r.stats.flat[:]
        for gi, ri, si in zip(gf, rf, sf):
            assert abs(ri.mean - gi.mean) < 0.03 * gi.sdev and abs(ri.sdev - gi.sdev) < 0.03 * gi.sdev
            assert abs(si.skew.mean) < max(5 * si.skew.sdev, 0.05) and abs(si.ex_kurt.mean) < max(5 * si.ex_kurt.sdev, 0.1)
            assert abs(gv.mean(si.median.loc) - gi.mean) < 0.1 * gi.sdev
            assert abs(gv.mean(si.median.plus) - gi.sdev) < 0.1 * gi.sdev
            assert abs(gv.mean(si.median.minus) - gi.sdev) < 0.1 * gi.sdev
        assert abs(gv.mean(np.asarray(r.vegas_mean, dtype=object).reshape(-1)[0] if not hasattr(r.vegas_mean, 'keys')
                           else r.vegas_mean.buf[0]) - gf[0].mean) < 0.05 * gf[0].sdev


def test_sample():
    """tests:1452-1486: weighted samples reproduce the means, widths and the correlation"""
    vegas, gv = _v()
    nbatch = 100000
    cov1 = np.array([[1., 0.99], [0.99, 1]])
    D = np.array([2, 1e-1])
    cov1 = D[None, :] * cov1 * D[:, None]
    g = gv.gvar([1, 2], cov1)
    pdf = vegas.PDFIntegrator(g, seed=7)
    pdf()
    for axis, mode in [(-1, 'rbatch'), (0, 'lbatch')]:
        w, p = pdf.sample(nbatch=nbatch, mode=mode)
        assert abs(np.sum(w) - 1) < 1e-12 and w.shape[0] >= nbatch
        wgts = w[:, None] if mode == 'lbatch' else w[None, :]
        pavg = np.sum(wgts * p, axis=axis)
        psdev = (np.sum(wgts * p ** 2, axis=axis) - pavg ** 2) ** 0.5
        assert str(gv.gvar(pavg, psdev)) == str(g)
        p0, p1 = (p[0], p[1]) if mode == 'rbatch' else (p[:, 0], p[:, 1])
        cov01 = float(np.sum(p0 * p1 * w) - pavg[0] * pavg[1])
        assert round(cov01, 2) == round(gv.evalcov(g)[0, 1], 2)
    gd = gv.BufferDict(a=g[0], b=g[1])
    pdf = vegas.PDFIntegrator(gd, seed=7)
    pdf()
    w, p = pdf.sample(nbatch=nbatch, mode='rbatch')
    pavg = dict(a=np.sum(w * p['a']), b=np.sum(w * p['b']))
    assert abs(pavg['a'] - 1) < 0.02 and abs(pavg['b'] - 2) < 0.002
    cov01 = np.sum(w * p['a'] * p['b']) - pavg['a'] * pavg['b']
    assert round(float(cov01), 2) == round(gv.evalcov(g)[0, 1], 2)


def test_device_integrands():
    """f(p) evaluated in HBM: a library functor (polynomial moments of a correlated Gaussian, known in closed
    form) and a @devicebatchintegrand; the host route gives the same numbers from the same samples"""
    import torch
    vegas, gv = _v()
    g = gv.gvar(*G2)
    fpoly = vegas.integrands.Poly(0.5, [1., 2.], [1, 2])          # 0.5 + p0 + 2 p1^2
    exact = 0.5 + 1. + 2. * (4. + 2. ** 2)
    gev = vegas.PDFIntegrator(g, seed=8, neval=20000)
    gev(nitn=5)
    r = gev(fpoly, nitn=5, adapt=False)
    assert abs(r.mean - exact) < 5 * r.sdev and r.sdev < 0.05

    @vegas.devicebatchintegrand
    def fdev(p):
        assert isinstance(p, torch.Tensor) and p.is_cuda
        return torch.stack([p[:, 0], p[:, 0] * p[:, 1]], dim=1)
    gev2 = vegas.PDFIntegrator(gev, seed=9)
    rd = gev2(fdev, nitn=3, adapt=False)
    assert abs(rd[0].mean - 1.) < 5 * rd[0].sdev and abs(rd[1].mean - (2. + .1)) < 5 * rd[1].sdev
    gev3 = vegas.PDFIntegrator(gev, seed=9)
    rh = gev3(vegas.lbatchintegrand(lambda p: np.stack([p[:, 0], p[:, 0] * p[:, 1]], axis=1)), nitn=3, adapt=False)
    np.testing.assert_allclose(gv.mean(rd), gv.mean(rh), rtol=1e-12)
    np.testing.assert_allclose(gv.sdev(rd), gv.sdev(rh), rtol=1e-9)


def test_save_extend_pickle_ravg():
    """tests:1730-1835: save / saveall, extend, pickling of results and integrator, ravg of PDFEV results"""
    vegas, gv = _v()
    g = gv.gvar(*G2)
    gev = vegas.PDFIntegrator(g, seed=10, neval=2000)
    gev(nitn=3)
    for f in (lambda p: p[0] + p[1], lambda p: [p[0], p[1]], lambda p: dict(a=p[0], b=[p[1], p[0] * p[1]])):
        buf, bufall = io.BytesIO(), io.BytesIO()
        r = gev(f, nitn=3, adapt=False, save=buf, saveall=bufall)
        assert gev.analyzer is None
        r1 = pickle.loads(pickle.dumps(r))
        assert type(r1) is type(r) and str(r1) == str(r)
        assert abs(gv.mean(r1.pdfnorm) - gv.mean(r.pdfnorm)) < 1e-15
        # streams hold one pickle per iteration; the last save() equals the returned result
        buf.seek(0)
        last = None
        while True:
            try:
                last = pickle.load(buf)
            except EOFError:
                break
        assert str(last) == str(r)
        bufall.seek(0)
        lastall = None
        while True:
            try:
                lastall = pickle.load(bufall)
            except EOFError:
                break
        r2, integ2 = lastall
        assert str(r2) == str(r) and isinstance(integ2, vegas.PDFIntegrator)
        np.testing.assert_allclose(integ2.map.grid, gev.map.grid, rtol=1e-15)
        # extend: a second run appended to the first equals one average over all six iterations
        rb = gev(f, nitn=3, adapt=False)
        n0 = len(r.itn_results)
        r.extend(rb)
        assert len(r.results.itn_results) == n0 + 3
        ru = vegas.ravg(rb, weighted=False)
        assert type(ru) is type(rb)
    integ3 = pickle.loads(pickle.dumps(gev))
    assert isinstance(integ3, vegas.PDFIntegrator) and integ3.scale == gev.scale and integ3.limit == gev.limit
    r3 = integ3(lambda p: p[0], nitn=2, adapt=False)
    assert abs(r3.mean - 1.) < 5 * r3.sdev
