"""Golden fixture for ``vegas.PDFIntegrator``: runs the UNMODIFIED reference -- the compiled ``_vegas``
module of oracle/_ref plus the reference's own ``src/vegas/__init__.py`` (linked, not copied, into a scratch
package directory), imported with the test-only gvar stand-in (whose ``PDF`` restates gvar's published
behaviour: principal axes of the correlation matrix with an svdcut) -- on a correlated 3-parameter Gaussian
with a dictionary-valued f(p).  Uniforms are injected through ``ran_array_generator`` (numpy default_rng).
Recorded: the tan-map grid of ``PDFIntegrator._make_map`` (from ``gvar.ranseed(1)``), the reference's own
integrand ``_f_lbatch`` on a fixed batch of theta, and per iteration the flat means / covariance of
``[pdf, f(p) pdf ...]``, ``sigf``, ``last_neval`` and the adapted grid.  Build container only:

    make -C oracle ref && python tests/golden/make_golden_pdf.py
"""
import functools
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INIT = '/root/reference/src/vegas/__init__.py'

pkg = tempfile.mkdtemp(prefix='refpkg_')
os.makedirs(os.path.join(pkg, 'vegas'))
refdir = os.path.join(ROOT, 'oracle', '_ref', 'vegas')
so = [n for n in os.listdir(refdir) if n.startswith('_vegas') and n.endswith('.so')][0]
os.symlink(os.path.join(refdir, so), os.path.join(pkg, 'vegas', so))
os.symlink(REF_INIT, os.path.join(pkg, 'vegas', '__init__.py'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gvar_shim'))
sys.path.insert(0, pkg)
sys.path.insert(0, ROOT)
import gvar            # noqa: E402  (the shim)
import vegas           # noqa: E402  (the reference package)

from tests.golden.cases import PDF_CASES, pdf_f      # noqa: E402


class Recorder(object):
    def __init__(self):
        self.rows = []

    def begin(self, itn, integ):
        self.integ = integ

    def end(self, itn_result, result):
        I = self.integ
        r = np.asarray(itn_result.buf, dtype=object).reshape(-1)
        self.rows.append(dict(mean=np.array([x.mean for x in r], float), cov=np.array(gvar.evalcov(r), float),
                              sigf=np.array(I.sigf, float), last_neval=int(I.last_neval), grid=np.array(I.map.grid, float)))


out = {}
for name, spec in PDF_CASES.items():
    gvar.ranseed(1)
    rng = np.random.default_rng(spec['seed'])
    g = gvar.gvar(spec['mean'], spec['cov'])
    rec = Recorder()
    integ = vegas.PDFIntegrator(g, scale=spec['scale'], limit=spec['limit'], adapt_to_pdf=spec['adapt_to_pdf'],
                                ran_array_generator=lambda shape: rng.random(shape), analyzer=rec, **spec['kw'])
    out[name + '_map0'] = np.array(integ.map.grid, float)
    out[name + '_vec_sig'] = np.array(integ.param_pdf.vec_sig, float)
    out[name + '_dp_dchiv'] = np.float64(integ.param_pdf.dp_dchiv)
    f = vegas.lbatchintegrand(pdf_f)
    # the reference's integrand on a fixed batch of theta
    th = np.random.default_rng(spec['seed'] + 1).uniform(-1.4, 1.4, size=(257, len(spec['mean'])))
    fstd = integ._make_std_integrand(f, integ.param_sample)
    ans = vegas.PDFIntegrator._f_lbatch(th, f=fstd, param_pdf=integ.param_pdf, pdf=None, scale=integ.scale,
                                        adapt_to_pdf=integ.adapt_to_pdf)
    out[name + '_theta'] = th
    out[name + '_keys'] = np.array([str(k) for k in ans.keys()])
    out[name + '_rows'] = np.concatenate([np.asarray(ans[k], float).reshape(len(th), -1) for k in ans], axis=1)
    r = integ(f, nitn=spec['nitn'])
    flat = np.asarray(r.buf if hasattr(r, 'buf') else r, dtype=object).reshape(-1)
    out[name + '_result_mean'] = np.array([x.mean for x in flat], float)
    out[name + '_result_sdev'] = np.array([x.sdev for x in flat], float)
    out[name + '_pdfnorm'] = np.array([r.pdfnorm.mean, r.pdfnorm.sdev])
    for i, row in enumerate(rec.rows):
        for k, v in row.items():
            out['%s_itn%d_%s' % (name, i, k)] = v
    print(name, 'map0[:3]', out[name + '_map0'][0, :3], 'result', r, 'pdfnorm', r.pdfnorm)
np.savez_compressed(os.path.join(HERE, 'ref_pdf.npz'), **out)
print('wrote ref_pdf.npz')
