"""``PDFIntegrator`` and its results ``PDFEV`` / ``PDFEVArray`` / ``PDFEVDict`` -- expectation values with
respect to a probability density (reference ``src/vegas/__init__.py:81-1217``), on the B200 engine.

What runs where.  The reference wraps the user's ``f(p)`` in a numpy lbatch integrand that maps the
integration variables ``theta`` to parameters ``p`` (a tan map along the principal axes of the
parameters' correlation matrix) and multiplies by the PDF.  Here that wrapper is a device batch
integrand: the samples ``theta[rows, dim]`` the sampler kernel wrote to HBM go through ``k_pdf_map``
(``csrc/pdfmap.cu``: ``theta -> p`` and the weight ``dp/dtheta * pdf`` in one pass), ``f(p)`` is
evaluated on the device when it is a ``@devicebatchintegrand`` or a ``DeviceIntegrand`` from the
library (else on host copies, like any numpy integrand), ``k_pdf_weight`` assembles the rows
``[pdf | f(p) pdf]`` in HBM, and the engine's reduce kernel takes it from there.  The numpy twin of the
wrapper (``PDFIntegrator._f_lbatch``, same formulas as the reference's) is used once per call to learn
the structure of the result, and by the tests as the kernels' oracle.
"""
import functools
import pickle

import numpy as np

from ._gv import gv
from ._integrand import DeviceIntegrand, VegasIntegrand, lbatchintegrand
from ._integrator import Integrator
from ._map import AdaptiveMap


# --------------------------------------------------------------------------- results
def _ratio(results):
    """<f(p)> from the integrals of f(p) pdf and pdf (src/vegas/__init__.py:139, 239, 328-332)"""
    return results['f(p)*pdf'] / results['pdf']


class _EVMixin(object):
    """what the three result types share: the underlying ``RAvgDict`` (``results``), its attributes by
    delegation, ``pdfnorm`` and ``extend``"""

    def _install(self, results, analyzer):
        object.__setattr__(self, 'results', results)
        object.__setattr__(self, 'analyzer', analyzer)

    @staticmethod
    def _load(results):
        return pickle.loads(results) if isinstance(results, bytes) else results

    def extend(self, pdfev):
        r""" Merge results from ``pdfev`` after the results currently in ``self``. """
        self.results.extend(pdfev.results)

    def _delegate(self, k):
        if k in ('keys', 'results', 'analyzer') or k.startswith('__'):
            raise AttributeError(k)
        if k == 'pdfnorm':
            return self.results['pdf']
        return getattr(self.results, k)


class PDFEV(_EVMixin, gv.GVar):
    r""" Expectation value (a |GVar|) from :class:`PDFIntegrator` (reference ``__init__.py:81-176``).
    Attributes: ``pdfnorm`` (integral of the PDF), ``results`` (the :class:`RAvgDict` of the underlying
    integrals, whose attributes -- ``Q``, ``chi2``, ``summary()`` ... -- are available here too); results
    of :meth:`PDFIntegrator.stats` also carry ``stats``, ``vegas_mean``, ``vegas_cov``, ``vegas_sdev``. """

    def __init__(self, results, analyzer=None):
        results = self._load(results)
        if analyzer is None:
            ans = _ratio(results)
        else:
            ans, extras = analyzer(results)
        gv.GVar.__init__(self, *ans.internaldata)
        self._install(results, analyzer)
        if analyzer is not None:
            for k in extras:
                object.__setattr__(self, k, extras[k])

    def __getattr__(self, k):
        return self._delegate(k)

    def __reduce_ex__(self, protocol):
        return (PDFEV, (pickle.dumps(self.results), self.analyzer))


class PDFEVArray(_EVMixin, np.ndarray):
    r""" Array of expectation values from :class:`PDFIntegrator` (reference ``__init__.py:178-272``). """

    def __new__(cls, results, analyzer=None):
        results = cls._load(results)
        if analyzer is None:
            self = np.asarray(_ratio(results)).view(cls)
            extras = {}
        else:
            ans, extras = analyzer(results)
            self = np.asarray(ans).view(cls)
        self._install(results, analyzer)
        for k in extras:
            object.__setattr__(self, k, extras[k])
        return self

    def __array_finalize__(self, obj):
        pass

    def __getattr__(self, k):
        return self._delegate(k)

    def __reduce_ex__(self, protocol):
        return (PDFEVArray, (pickle.dumps(self.results), self.analyzer))


class PDFEVDict(_EVMixin, gv.BufferDict):
    r""" Dictionary of expectation values from :class:`PDFIntegrator` (reference ``__init__.py:274-371``). """

    def __init__(self, results, analyzer=None):
        gv.BufferDict.__init__(self)
        results = self._load(results)
        self._install(results, analyzer)
        if analyzer is None:
            for k in results:
                if k != 'pdf':
                    self[k[1]] = results[k]
            self.buf[:] /= results['pdf']
        else:
            ans, extras = analyzer(results)
            for k in extras:
                object.__setattr__(self, k, extras[k])
            for k in ans:
                self[k] = ans[k]

    def __getattr__(self, k):
        return self._delegate(k)

    def __reduce_ex__(self, protocol):
        return (PDFEVDict, (pickle.dumps(self.results), self.analyzer))


# --------------------------------------------------------------------------- the integrand on the device
class _DevicePDFIntegrand(object):
    """``theta[rows, dim]`` (CUDA tensor) -> rows ``[pdf | f(p) pdf]`` (or ``[f(p) pdf | pdf]``) in HBM"""

    def __init__(self, integ, f, fdev):
        import torch
        self.torch = torch
        self.integ = integ
        self.f = f                       # VegasIntegrand over the parameter layout, or None
        self.fdev = fdev                 # device twin of a library functor, or None
        ctx, _ = integ._engine()
        self.ctx = ctx
        pdf = integ.param_pdf
        self.mean = torch.from_numpy(np.ascontiguousarray(pdf.meanflat, dtype=float)).to(ctx.device)
        self.vec_sig = torch.from_numpy(np.ascontiguousarray(pdf.vec_sig, dtype=float)).to(ctx.device)

    def _on_host(self, std, p):
        out = np.asarray(std.eval(p.cpu().numpy(), jac=None), dtype=float).reshape(p.shape[0], -1)
        return self.torch.from_numpy(np.ascontiguousarray(out)).to(p.device)

    def __call__(self, theta, jac=None):
        torch, integ = self.torch, self.integ
        rows = theta.shape[0]
        p = torch.empty_like(theta)
        w = torch.empty(rows, dtype=torch.float64, device=theta.device)
        self.ctx.pdf_map(theta, integ.scale, integ.param_pdf.dp_dchiv, integ.pdf is None, self.mean, self.vec_sig, p, w)
        if integ.pdf is not None:        # the user's PDF: product of its components (__init__.py:609)
            pv = integ.pdf.eval(p) if integ.pdf.on_device else self._on_host(integ.pdf, p)
            w = w * pv.reshape(rows, -1).prod(dim=1)
        if self.f is None:
            return w.reshape(rows, 1)
        if self.fdev is not None:
            fp = self.fdev(p)
        elif self.f.on_device:
            fp = self.f.eval(p)
        else:
            fp = self._on_host(self.f, p)
        fp = fp.reshape(rows, -1).contiguous()
        out = torch.empty((rows, fp.shape[1] + 1), dtype=torch.float64, device=theta.device)
        self.ctx.pdf_weight(fp, w, out, integ.adapt_to_pdf)
        return out


# --------------------------------------------------------------------------- the integrator
class PDFIntegrator(Integrator):
    r""" :mod:`vegas` integrator for PDF expectation values (reference ``__init__.py:373-1188``).

    ``PDFIntegrator(param, pdf)`` evaluates expectation values of functions ``f(p)`` with respect to the
    probability density ``pdf(p)`` (default: the Gaussian distribution of ``param``).  ``param`` -- a
    |GVar|, an array of them, or a dictionary -- defines the integration variables: the parameters are
    re-expressed along the principal axes of ``param``'s correlation matrix and mapped to a finite
    range by ``p = mean + scale * tan(theta)`` in units of the standard deviations, out to ``limit``
    standard deviations.  ``adapt_to_pdf`` (default ``True``) makes |vegas| adapt to the PDF rather than to
    ``f(p) * pdf(p)``; ``svdcut`` regulates small eigenvalues of the correlation matrix.  All other
    keywords go to :class:`Integrator`; ``uses_jac`` is ignored.

    ``g_ev(f)`` returns ``<f(p)>`` as :class:`PDFEV`, :class:`PDFEVArray` or :class:`PDFEVDict` (``f`` may
    return a number, an array or a dictionary; batch integrands, ``@devicebatchintegrand`` and library
    ``DeviceIntegrand`` functors are evaluated on batches); ``result.pdfnorm`` is the integral of the PDF.
    """

    def __init__(self, param=None, pdf=None, adapt_to_pdf=True, limit=100., scale=1., svdcut=1e-15, **kargs):
        if 'g' in kargs and param is None:          # legacy name
            kargs = dict(kargs)
            param = kargs.pop('g')
        if param is None:
            raise ValueError('param must be specified')
        if isinstance(param, PDFIntegrator):
            super(PDFIntegrator, self).__init__(param, **{k: v for k, v in kargs.items() if k != 'uses_jac'})
            for k in ['param_pdf', 'param_sample', 'pdf', 'adapt_to_pdf', 'limit', 'scale']:
                setattr(self, k, getattr(param, k))
            return
        self.param_pdf = param if isinstance(param, gv.PDF) else gv.PDF(param, svdcut=svdcut)
        self.param_sample = self.param_pdf.sample(mode=None)
        self.limit = abs(limit)
        self.scale = abs(scale)
        self.set(adapt_to_pdf=adapt_to_pdf, pdf=pdf)
        kargs = {k: v for k, v in kargs.items() if k != 'uses_jac'}
        device = kargs.get('device', None)
        integ_map = self._make_map(self.limit / self.scale, device)
        super(PDFIntegrator, self).__init__(AdaptiveMap(self.param_pdf.size * [integ_map]), **kargs)

    def __reduce__(self):
        kargs = dict()
        for k in Integrator.defaults:
            if k not in ('uses_jac', 'map') and not _same(Integrator.defaults[k], getattr(self, k)):
                kargs[k] = getattr(self, k)
        for k in Integrator.engine_defaults:
            if k != 'device':
                kargs[k] = getattr(self, k)
        kargs['map'] = self.map
        kargs['nstrat'] = np.asarray(self.nstrat)
        kargs['sigf'] = np.array(self.sigf)
        kargs['_itn_counter'] = self._itn_counter
        return (PDFIntegrator, (self.param_pdf, self.pdf, self.adapt_to_pdf, self.limit, self.scale), kargs)

    def __setstate__(self, kargs):
        kargs = dict(kargs)
        self._itn_counter = int(kargs.pop('_itn_counter', 0))
        engine = {k: kargs.pop(k) for k in list(kargs) if k in Integrator.engine_defaults}
        for k, v in engine.items():
            setattr(self, k, v)
        self.set(**kargs)

    def set(self, ka={}, **kargs):
        r""" Reset default parameters of the integrator (``pdf`` and ``adapt_to_pdf`` included; ``param``
        cannot be changed).  Returns the old values, as :meth:`Integrator.set` does. """
        if kargs:
            kargs.update(ka)
        else:
            kargs = dict(ka)
        old = {}
        if 'param' in kargs:
            raise ValueError("Can't reset param.")
        if 'pdf' in kargs:
            if hasattr(self, 'pdf'):
                old['pdf'] = self.pdf
            pdf = kargs.pop('pdf')
            self.pdf = pdf if pdf is None else self._make_std_integrand(pdf, xsample=self.param_sample)
        if 'adapt_to_pdf' in kargs:
            if hasattr(self, 'adapt_to_pdf'):
                old['adapt_to_pdf'] = self.adapt_to_pdf
            self.adapt_to_pdf = kargs.pop('adapt_to_pdf')
        if kargs:
            old.update(super(PDFIntegrator, self).set(kargs))
        return old

    def _make_std_integrand(self, fcn, xsample=None):
        if isinstance(fcn, VegasIntegrand):
            return fcn
        if not hasattr(self, 'map'):                 # set(pdf=...) in __init__, before Integrator.__init__
            return VegasIntegrand(fcn=fcn, map=None, uses_jac=False, xsample=xsample, mpi=False)
        return super(PDFIntegrator, self)._make_std_integrand(fcn, xsample=xsample)

    def _make_map(self, limit, device=None):
        r""" One-dimensional grid adapted to a unit Gaussian in ``scale * tan(theta)`` (``__init__.py:576-591``):
        ten adaptations of a 100-increment map on 2000 random points from ``gvar.RNG``. """
        ny = 2000
        y = gv.RNG.random((ny, 1))
        limit = np.arctan(limit)
        m = AdaptiveMap([[-limit, limit]], ninc=100)
        theta = np.empty(y.shape, float)
        jac = np.empty(y.shape[0], float)
        for _ in range(10):
            m.map(y, theta, jac)
            tan_theta = np.tan(theta[:, 0])
            x = self.scale * tan_theta
            fx = (tan_theta ** 2 + 1) * np.exp(-(x ** 2) / 2.)
            m.add_training_data(y, (jac * fx) ** 2)
            m.adapt(alpha=1.5)
        return np.array(m.grid[0])

    @staticmethod
    def _f_lbatch(theta, f, param_pdf, pdf, scale, adapt_to_pdf):
        r""" The integrand in numpy (``__init__.py:593-640``): ``theta[i, d]`` -> dictionary with ``'pdf'`` and
        ``'f(p)*pdf'`` (or ``('f(p)*pdf', k)`` for every key ``k`` of a dictionary-valued ``f``).  The device
        path computes the same rows; this twin tells :class:`VegasIntegrand` their structure. """
        tan_theta = np.tan(theta)
        chiv = scale * tan_theta
        dp_dtheta = np.prod(scale * (tan_theta ** 2 + 1.), axis=1) * param_pdf.dp_dchiv
        p = param_pdf.pflat(chiv, mode='lbatch')
        if pdf is None:
            # normalized in chiv space, so param_pdf.dp_dchiv must not be in the Jacobian
            pdfv = np.prod(np.exp(-(chiv ** 2) / 2.) / np.sqrt(2 * np.pi), axis=1) / param_pdf.dp_dchiv
        else:
            pdfv = np.prod(_eval_on_host(pdf, p).reshape(p.shape[0], -1), axis=1)
        ans = gv.BufferDict()
        wgt = dp_dtheta * pdfv
        if f is None:
            ans['pdf'] = wgt
            return ans
        fp = f.format_evalx(_eval_on_host(f, p))
        if adapt_to_pdf:
            ans['pdf'] = wgt
        if hasattr(fp, 'keys'):
            for k in fp:
                fk = np.asarray(fp[k], dtype=float)
                ans[('f(p)*pdf', k)] = fk * wgt.reshape(fk.shape[:1] + (fk.ndim - 1) * (1,))
        else:
            fp = np.asarray(fp, dtype=float)
            ans['f(p)*pdf'] = fp * wgt.reshape(fp.shape[:1] + (fp.ndim - 1) * (1,))
        if not adapt_to_pdf:
            ans['pdf'] = wgt
        return ans

    def __call__(self, f=None, save=None, saveall=None, **kargs):
        r""" Estimate the expectation value of ``f(p)`` (``__init__.py:642-741``): integrates ``f(p) * pdf(p)``
        and ``pdf(p)`` together and returns their ratio(s) as :class:`PDFEV`, :class:`PDFEVArray` or
        :class:`PDFEVDict`.  ``f=None`` integrates only the PDF and returns the :class:`RAvgDict` with
        ``pdfnorm`` set.  ``pdf=...`` / ``adapt_to_pdf=...`` and all :class:`Integrator` keywords may be given
        here too; ``save`` / ``saveall`` pickle the result (and the integrator) after every iteration. """
        kargs = {k: v for k, v in kargs.items() if k != 'uses_jac'}
        if kargs:
            self.set(kargs)
        if save is not None or saveall is not None:
            self.set(analyzer=PDFAnalyzer(self, analyzer=self.analyzer, save=save, saveall=saveall))
        fdev = None
        if f is not None:
            if isinstance(f, DeviceIntegrand) and self.param_pdf.size <= 20:
                fdev = f.device_twin(self.param_pdf.size, device=self._engine()[0].device)
            f = self._make_std_integrand(f, self.param_sample)
        twin = lbatchintegrand(functools.partial(
            PDFIntegrator._f_lbatch, f=f, param_pdf=self.param_pdf, pdf=self.pdf, scale=self.scale,
            adapt_to_pdf=self.adapt_to_pdf))
        std = super(PDFIntegrator, self)._make_std_integrand(twin)          # one probe call: structure of the result
        std.eval = _DevicePDFIntegrand(self, f, fdev)
        std.on_device = True
        try:
            results = super(PDFIntegrator, self).__call__(std)
        finally:
            if isinstance(self.analyzer, PDFAnalyzer):
                self.set(analyzer=self.analyzer.analyzer)
        if gv.mean(results['pdf']) == 0:
            raise RuntimeError('Integral of PDF vanishes; increase neval?')
        if f is None:
            results.pdfnorm = results['pdf']
            return results
        return PDFIntegrator._make_ans(results)

    @staticmethod
    def _make_ans(results):
        if 'f(p)*pdf' not in results:
            return PDFEVDict(results)
        if np.ndim(results['f(p)*pdf']) == 0:
            return PDFEV(results)
        return PDFEVArray(results)

    # ---------------------------------------------------------------- statistics of f(p)
    def stats(self, f=None, moments=False, histograms=False, **kargs):
        r""" Statistical analysis of ``f(p)`` (``__init__.py:743-938``): means and (co)variances of the
        components of ``f(p)`` with respect to the PDF, returned as |GVar|\s whose standard deviations are
        those of the distribution (plus the |vegas| errors in quadrature).  The result also has
        ``stats`` (:class:`gvar.PDFStatistics` per component; with ``moments=True`` skewness and excess
        kurtosis, with ``histograms=True`` -- or a dictionary with ``nbin``, ``binwidth``, ``loc`` -- the
        histograms), ``vegas_mean``, ``vegas_cov`` and ``vegas_sdev``.  Adaptation is off (``adapt=False``)
        unless asked for. """
        oldsettings = {}
        if 'adapt' not in kargs:
            oldsettings['adapt'] = self.adapt
            kargs['adapt'] = False
        if f is None:
            f = lbatchintegrand(_identity)
        f = self._make_std_integrand(f, xsample=self.param_sample)
        fpsample = f(self.param_sample)
        if histograms is not False:
            histograms = {} if histograms is True else dict(histograms)
            nbin = histograms.setdefault('nbin', 12)
            binwidth = histograms.setdefault('binwidth', 0.5)
            loc = histograms.get('loc', None)
            if loc is not None:
                loc = gv.asbufferdict(loc).buf if hasattr(loc, 'keys') else np.asarray(loc).reshape(-1)
                mean, sdev = np.asarray(gv.mean(loc), float).reshape(-1), np.asarray(gv.sdev(loc), float).reshape(-1)
            else:
                # one iteration to locate the distributions of the components of f(p)
                oldnitn = self.nitn
                r = self(lbatchintegrand(functools.partial(_f_f2, f=f)), nitn=1)
                self.set(nitn=oldnitn)
                mean = np.asarray(gv.mean(r['f']), float).reshape(-1)
                sdev = np.fabs(np.asarray(gv.mean(r['f2']), float).reshape(-1) - mean * mean) ** 0.5
            halfwidth = nbin / 2 * binwidth
            histograms['bins'] = np.array([
                mean[i] + np.linspace(-halfwidth * sdev[i], halfwidth * sdev[i], nbin + 1) for i in range(mean.shape[0])])
        integrand = lbatchintegrand(functools.partial(
            PDFIntegrator._stats_integrand, f=f, moments=moments, histograms=histograms))
        integrand = self._make_std_integrand(integrand, xsample=np.asarray(gv.mean(_flat(self.param_sample)), float))
        results = self(integrand, **kargs)
        analyzer = functools.partial(
            PDFIntegrator._stats_analyzer, fpsample=fpsample, moments=moments, histograms=histograms)
        if getattr(fpsample, 'shape', ()) is None:
            ans = PDFEVDict(results.results, analyzer)
        elif np.shape(fpsample) == ():
            ans = PDFEV(results.results, analyzer)
        else:
            ans = PDFEVArray(results.results, analyzer)
        if oldsettings:
            self.set(**oldsettings)
        return ans

    @staticmethod
    def _stats_integrand(p, f, moments=False, histograms=False):
        r""" ``f(p)``, the products of its components and, on request, third and fourth powers and histogram
        counts (``__init__.py:1019-1041``); ``p[i, d]`` are flat parameter values """
        fp = np.asarray(f.eval(p), dtype=float)
        nbatch, nfp = fp.shape
        iu, ju = np.tril_indices(nfp)
        ans = gv.BufferDict()
        ans['fp'] = fp
        ans['fpfp'] = fp[:, iu] * fp[:, ju]                 # row-major lower triangle: (0,0), (1,0), (1,1), ...
        if moments:
            ans['fp**3'] = fp ** 3
            ans['fp**4'] = fp ** 4
        if histograms:
            count = np.zeros((nbatch, nfp, histograms['nbin'] + 2), dtype=float)
            idx = np.arange(nbatch)
            for j in range(nfp):
                count[idx, j, np.searchsorted(histograms['bins'][j], fp[:, j], side='right')] = 1
            ans['count'] = count
        return ans

    @staticmethod
    def _stats_analyzer(results, fpsample, moments, histograms):
        r""" Final :meth:`stats` results from the integrals (``__init__.py:940-1017``) """
        pdfnorm = results['pdf']
        ev = {k[1]: results[k] / pdfnorm for k in results if k != 'pdf'}
        fp = np.asarray(ev['fp'], dtype=object).reshape(-1)
        nfp = fp.shape[0]
        meanfp = np.asarray(gv.mean(fp), float)
        covfpfp = np.zeros((nfp, nfp), dtype=object)
        fp2 = np.empty(nfp, dtype=object)
        fpfp = iter(np.asarray(ev['fpfp'], dtype=object).reshape(-1))
        for i in range(nfp):
            for j in range(i + 1):
                v = next(fpfp)
                if i == j:
                    fp2[i] = v
                    covfpfp[i, i] = v - fp[i] ** 2
                else:
                    covfpfp[i, j] = covfpfp[j, i] = v - fp[i] * fp[j]
        # |vegas| errors added to the distribution's covariance
        ans = gv.gvar(meanfp, np.asarray(gv.mean(covfpfp), float) + np.asarray(gv.evalcov(fp), float).reshape(nfp, nfp))
        shape = getattr(fpsample, 'shape', ())
        if shape is None:
            ans = gv.BufferDict(fpsample, buf=ans)
            mean = gv.BufferDict(fpsample, buf=fp)
            cov, sdev = gv.BufferDict(), gv.BufferDict()
            for k in mean:
                ksl, kshape = _as_slice(mean.slice(k)), np.shape(mean[k])
                for l in mean:
                    lsl, lshape = _as_slice(mean.slice(l)), np.shape(mean[l])
                    block = covfpfp[ksl, lsl]
                    cov[k, l] = block.reshape(kshape + lshape) if kshape + lshape != () else block[0, 0]
                d = gv.fabs(np.diag(covfpfp[ksl, ksl])) ** 0.5
                sdev[k] = d[0] if kshape == () else d.reshape(kshape)
        elif np.shape(fpsample) == ():
            ans, mean, cov = ans.flat[0], fp.flat[0], covfpfp
            sdev = gv.fabs(cov) ** 0.5
        else:
            shape = np.shape(fpsample)
            ans, mean = ans.reshape(shape), fp.reshape(shape)
            cov = covfpfp.reshape(shape + shape)
            sdev = (gv.fabs(np.diag(covfpfp)) ** 0.5).reshape(shape)
        stats = np.empty(nfp, dtype=object)
        for i in range(nfp):
            mom = [fp[i], fp2[i]]
            if moments:
                mom += [np.asarray(ev['fp**3'], dtype=object).reshape(-1)[i], np.asarray(ev['fp**4'], dtype=object).reshape(-1)[i]]
            hist = (histograms['bins'][i], np.asarray(ev['count'], dtype=object).reshape(nfp, -1)[i]) if histograms else None
            stats[i] = gv.PDFStatistics(moments=mom, histogram=hist)
        if getattr(fpsample, 'shape', ()) is None:
            stats = gv.BufferDict(fpsample, buf=stats)
        elif np.shape(fpsample) == ():
            stats = stats.flat[0]
        else:
            stats = stats.reshape(np.shape(fpsample))
        return ans, dict(stats=stats, vegas_mean=mean, vegas_cov=cov, vegas_sdev=sdev)

    # ---------------------------------------------------------------- samples from the PDF
    def sample(self, nbatch, mode='rbatch'):
        r""" Weighted random samples from the integrator's PDF (``__init__.py:1043-1188``): ``wgts, samples``
        with ``sum(wgts) == 1`` and at least ``nbatch`` samples (a multiple of ``last_neval``), laid out like
        ``param`` plus a batch index on the right (``mode='rbatch'``) or on the left (``'lbatch'``).  The
        ``theta -> p`` map and the weights come from ``k_pdf_map``, batch by batch on the device. """
        import torch
        neval = self.last_neval if getattr(self, 'last_neval', 0) > 0 else self.neval
        nit = 1 if nbatch is None else int(nbatch) // int(neval)
        if nbatch is not None and nit * neval < nbatch:
            nit += 1
        dev = _DevicePDFIntegrand(self, None, None)
        samples, wgts = [], []
        for _ in range(nit):
            for theta, wgt in self.random_batch_device():
                rows = theta.shape[0]
                p = torch.empty_like(theta)
                w = torch.empty(rows, dtype=torch.float64, device=theta.device)
                dev.ctx.pdf_map(theta, self.scale, self.param_pdf.dp_dchiv, self.pdf is None, dev.mean, dev.vec_sig, p, w)
                if self.pdf is not None:
                    pv = self.pdf.eval(p) if self.pdf.on_device else dev._on_host(self.pdf, p)
                    w = w * pv.reshape(rows, -1).prod(dim=1)
                wgts.append((wgt * w).cpu().numpy())
                samples.append(p.cpu().numpy())
        samples = np.concatenate(samples, axis=0)
        wgts = np.concatenate(wgts)
        wgts /= np.sum(wgts)
        if mode == 'rbatch':
            return wgts, self.param_pdf._unflatten(samples.T, mode='rbatch')
        return wgts, self.param_pdf._unflatten(samples, mode='lbatch')


class PDFAnalyzer(object):
    r""" |vegas| analyzer implementing the ``save`` / ``saveall`` keywords of :class:`PDFIntegrator`
    (reference ``__init__.py:1190-1217``) """

    def __init__(self, pdfinteg, analyzer, save=None, saveall=None):
        self.pdfinteg, self.analyzer, self.save, self.saveall = pdfinteg, analyzer, save, saveall

    def begin(self, itn, integrator):
        if self.analyzer is not None:
            self.analyzer.begin(itn, integrator)

    def end(self, itn_result, results):
        if self.analyzer is not None:
            self.analyzer.end(itn_result, results)
        if self.save is None and self.saveall is None:
            return
        ans = PDFIntegrator._make_ans(results) if len(list(results.keys())) > 1 else results
        rank, world = self.pdfinteg._rank_world()
        # (pickling the integrator gathers sigf from every rank: all ranks pickle, rank 0 writes)
        for target, obj in ((self.save, ans), (self.saveall, (ans, self.pdfinteg))):
            if target is None:
                continue
            payload = pickle.dumps(obj)
            if world > 1 and rank != 0:
                continue
            if isinstance(target, str):
                with open(target, 'wb') as ofile:
                    ofile.write(payload)
            else:
                target.write(payload)


def _eval_on_host(std, p):
    """values of a standard-form integrand at host points ``p[i, d]`` (device integrands: through HBM and back)"""
    if getattr(std, 'on_device', False):
        import torch
        out = std.eval(torch.from_numpy(np.ascontiguousarray(p, dtype=float)).cuda())
        return out.cpu().numpy().reshape(p.shape[0], -1)
    return np.asarray(std.eval(p), dtype=float)


def _identity(p, jac=None):
    return p


def _f_f2(p, f):
    if hasattr(p, 'keys'):
        lb = getattr(p, 'lbatch_buf', None)
        p = np.asarray(lb if lb is not None else p.buf, dtype=float)
    else:
        p = np.reshape(p, (np.shape(p)[0], -1))
    fp = np.asarray(f.eval(p), dtype=float)
    return dict(f=fp, f2=fp ** 2)


def _as_slice(sl):
    return sl if isinstance(sl, slice) else slice(sl, sl + 1)


def _flat(sample):
    return sample.buf if hasattr(sample, 'keys') else np.asarray(sample).reshape(-1)


def _same(a, b):
    try:
        return bool(np.all(a == b))
    except Exception:
        return a is b
