"""CUDA engine against the REFERENCE's own recorded outputs -- no oracle in between.

tests/golden/ref_*.npz hold what the unmodified reference (``/root/reference/src/vegas/_vegas.pyx``
compiled by ``oracle/Makefile``; generator script ``tests/golden/make_golden.py``) computed for a few
iterations of 13 configurations when fed a recorded uniform stream through its documented injection
hook ``Integrator.ran_array_generator`` (pyx:1081-1086, 1676-1680, 1732).  Here the same stream goes
into ``vegas_b200.Integrator(..., ran_array_generator=...)``: the CUDA sampler consumes the injected
uniforms (``vb200_sample_from_uniforms``), the numpy integrand runs on host copies of the GPU samples,
and the CUDA reduce kernel produces the iteration's sums.  Integers must be exact; floating-point
results agree to the tolerances of ``north_star`` (1e-12 relative; covariances of few-sample cubes
1e-11)."""
import os

import numpy as np
import pytest

import vegas_b200 as vegas
from tests.golden.cases import CASES, integrand

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


class Recorder(object):
    """analyzer hook (pyx:1009-1026): snapshot of the integrator after every iteration"""

    def __init__(self):
        self.rows = []

    def begin(self, itn, integ):
        self.integ = integ

    def end(self, itn_result, result):
        I = self.integ
        self.rows.append(dict(sigf=np.array(I.sigf, float), sum_sigf=float(I.sum_sigf), last_neval=int(I.last_neval),
                              range=np.array(I.neval_hcube_range, np.int64), grid=np.array(I.map.grid, float)))


@pytest.mark.parametrize('name', sorted(CASES))
def test_cuda_replays_reference_fixture(name):
    spec = CASES[name]
    G = np.load(os.path.join(HERE, 'golden', 'ref_%s.npz' % name))
    rng = np.random.default_rng(spec['seed'])
    rows = []

    def gen(shape):
        rows.append(int(shape[0]))
        return rng.random(shape)

    rec = Recorder()
    sums = []
    integ = vegas.Integrator(spec['limits'], ran_array_generator=gen, analyzer=rec, **spec['kw'])
    integ._trace = sums.append
    assert list(integ.nstrat) == list(G['nstrat']) and list(integ.map.ninc) == list(G['ninc'])
    assert integ.nhcube == int(G['nhcube']) and integ.min_neval_hcube == int(G['min_neval_hcube'])
    integ(vegas.lbatchintegrand(integrand(spec['f'])), nitn=spec['nitn'])
    assert len(rec.rows) == spec['nitn']
    adaptive = spec['kw'].get('beta', 0.75) > 0 and not spec['kw'].get('adapt_to_errors', False)
    for i, (row, raw) in enumerate(zip(rec.rows, sums)):
        assert row['last_neval'] == int(G['itn%d_last_neval' % i]), i
        np.testing.assert_allclose(raw['mean'], G['itn%d_mean' % i], rtol=1e-12, atol=1e-300)
        cov, var = G['itn%d_cov' % i], raw['var']
        if np.ndim(var) == 2:
            np.testing.assert_allclose(var, cov, rtol=1e-11, atol=1e-18 * np.abs(cov).max())
        else:
            np.testing.assert_allclose(var, np.diag(cov), rtol=1e-11)
        if len(G['itn%d_sigf' % i]) and adaptive and spec['kw'].get('adapt', True):
            np.testing.assert_allclose(row['sigf'], G['itn%d_sigf' % i], rtol=1e-8, atol=1e-300)
            np.testing.assert_allclose(row['sum_sigf'], float(G['itn%d_sum_sigf' % i]), rtol=1e-12)
        if adaptive:
            assert tuple(row['range']) == tuple(G['itn%d_range' % i]), i
        g = G['itn%d_grid' % i]
        for d in range(integ.dim):
            n = integ.map.ninc[d] + 1
            np.testing.assert_allclose(row['grid'][d, :n], g[d, :n], rtol=1e-12, atol=1e-15)
    # every uniform of the recorded stream was consumed, in the same order (batches are cut at other
    # places than the reference's -- ours end on 256-cube chunks -- which a sequential stream does not see)
    assert sum(rows) == int(np.sum(G['batch_rows']))


def test_ran_array_generator_feeds_random_batch():
    """``random_batch`` draws from the injected generator too (pyx:1732): x = map(y(u))"""
    calls = []

    def gen(shape):
        calls.append(shape)
        return np.full(shape, 0.5)

    integ = vegas.Integrator([[0., 1.], [0., 2.]], neval=200, ran_array_generator=gen)
    xs = np.concatenate([x for x, w in integ.random_batch()])
    assert calls and sum(s[0] for s in calls) == len(xs)
    ns = np.asarray(integ.nstrat)
    # u = 1/2 puts every sample at the centre of its stratum on the (still uniform) grid
    centres0 = (np.arange(ns[0]) + 0.5) / ns[0]
    assert np.allclose(np.unique(np.round(xs[:, 0], 12)), np.round(centres0, 12))


# ---------------------------------------------------------------------------------------------
# PDFIntegrator against the unmodified reference (tests/golden/make_golden_pdf.py -> ref_pdf.npz)
# ---------------------------------------------------------------------------------------------
from tests.golden.cases import PDF_CASES, pdf_f      # noqa: E402


@pytest.mark.parametrize('name', sorted(PDF_CASES))
def test_pdfintegrator_replays_reference_fixture(name):
    """the reference's ``PDFIntegrator`` recorded: its tan-map grid (``_make_map`` from ``gvar.ranseed(1)``), its
    integrand ``_f_lbatch`` on a fixed batch of theta, and three iterations on an injected uniform stream.
    Ours: the same construction (map adapted by the CUDA map kernels), ``k_pdf_map`` / ``k_pdf_weight`` on
    the same theta, and the same stream through sampler -> device wrapper -> reduce."""
    import torch
    from vegas_b200._gv import gv
    from vegas_b200._pdf import _DevicePDFIntegrand
    spec = PDF_CASES[name]
    G = np.load(os.path.join(HERE, 'golden', 'ref_pdf.npz'))
    gv.ranseed(1)
    rng = np.random.default_rng(spec['seed'])
    rec = Recorder()
    sums = []
    g = gv.gvar(spec['mean'], spec['cov'])
    integ = vegas.PDFIntegrator(g, scale=spec['scale'], limit=spec['limit'], adapt_to_pdf=spec['adapt_to_pdf'],
                                ran_array_generator=lambda shape: rng.random(shape), analyzer=rec, **spec['kw'])
    integ._trace = sums.append
    np.testing.assert_allclose(integ.param_pdf.vec_sig, G[name + '_vec_sig'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(integ.param_pdf.dp_dchiv, float(G[name + '_dp_dchiv']), rtol=1e-13)
    # ten adaptations of the 1-d map on 2000 points: CUDA map / add_training_data + host adapt vs the reference's
    np.testing.assert_allclose(integ.map.grid, G[name + '_map0'], rtol=1e-10, atol=1e-12)
    # the integrand on the recorded theta
    f = vegas.lbatchintegrand(pdf_f)
    fstd = integ._make_std_integrand(f, integ.param_sample)
    dev = _DevicePDFIntegrand(integ, fstd, None)
    rows = dev(torch.from_numpy(G[name + '_theta']).cuda()).cpu().numpy()
    assert list(G[name + '_keys']) == (["pdf", "('f(p)*pdf', 'a')", "('f(p)*pdf', 'b')"] if spec['adapt_to_pdf']
                                       else ["('f(p)*pdf', 'a')", "('f(p)*pdf', 'b')", "pdf"])
    np.testing.assert_allclose(rows, G[name + '_rows'], rtol=1e-12, atol=1e-300)
    # the iterations, from the reference's own initial map
    integ.set(map=vegas.AdaptiveMap(G[name + '_map0']))
    r = integ(f, nitn=spec['nitn'])
    assert len(rec.rows) == spec['nitn']
    for i, (row, raw) in enumerate(zip(rec.rows, sums)):
        assert row['last_neval'] == int(G['%s_itn%d_last_neval' % (name, i)]), i
        np.testing.assert_allclose(raw['mean'], G['%s_itn%d_mean' % (name, i)], rtol=1e-11, atol=1e-300)
        cov = G['%s_itn%d_cov' % (name, i)]
        np.testing.assert_allclose(raw['var'], cov, rtol=1e-9, atol=1e-16 * np.abs(cov).max())
        np.testing.assert_allclose(row['sigf'], G['%s_itn%d_sigf' % (name, i)], rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(row['grid'], G['%s_itn%d_grid' % (name, i)], rtol=1e-10, atol=1e-13)
    flat = np.asarray(r.buf, dtype=object).reshape(-1)
    np.testing.assert_allclose([x.mean for x in flat], G[name + '_result_mean'], rtol=1e-9)
    np.testing.assert_allclose([x.sdev for x in flat], G[name + '_result_sdev'], rtol=1e-6)
    np.testing.assert_allclose([r.pdfnorm.mean, r.pdfnorm.sdev], G[name + '_pdfnorm'], rtol=1e-6)
