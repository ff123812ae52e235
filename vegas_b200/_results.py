"""Result accumulators ``RAvg`` / ``RAvgArray`` / ``RAvgDict`` and ``VegasResult`` -- the running
(inverse-variance weighted, or plain) averages over iterations with chi2 / dof / Q
(reference ``_vegas.pyx:2271-2957``).  A few numbers per iteration: host arithmetic only.
"""
import pickle
import sys
import time

import numpy as np

from ._gv import gv
from ._map import TINY, EPSILON


def _summary_table(itn_results, make_acc, first, weighted):
    """iteration-by-iteration table shared by the three result classes"""
    acc = make_acc()
    linedata = []
    for i, res in enumerate(itn_results):
        acc.add(res)
        itn = '%3d' % (i + 1)
        integral = '%-15s' % first(res)
        wgtavg = '%-15s' % first(acc)
        chi2dof = '%8.2f' % (acc.chi2 / acc.dof if i != 0 else 0.0)
        Q = '%8.2f' % (acc.Q if i != 0 else 1.0)
        linedata.append((itn, integral, wgtavg, chi2dof, Q))
    nchar = 5 * [0]
    for data in linedata:
        for i, d in enumerate(data):
            nchar[i] = max(nchar[i], len(d))
    fmt = '%%%ds   %%-%ds %%-%ds %%%ds %%%ds\n' % tuple(nchar)
    ans = fmt % ('itn', 'integral', 'wgt average' if weighted else 'average', 'chi2/dof', 'Q')
    ans += len(ans[:-1]) * '-' + '\n'
    for data in linedata:
        ans += fmt % data
    return ans


class RAvg(gv.GVar):
    r""" Running average of scalar-valued Monte Carlo estimates (``_vegas.pyx:2276-2451``).

    Estimates are weighted by their inverse variances if ``weighted=True``; otherwise straight,
    unweighted averages are used. """

    def __init__(self, weighted=True, itn_results=None, sum_neval=0, _rescale=True):
        self.rescale = None
        if weighted:
            self._wlist = []
            self.weighted = True
        else:
            self._msum = 0.
            self._varsum = 0.
            self._n = 0
            self.weighted = False
        self._mlist = []
        self.itn_results = []
        if itn_results is None:
            super(RAvg, self).__init__(*gv.gvar(0., 0.).internaldata)
        else:
            if isinstance(itn_results, bytes):
                itn_results = gv.loads(itn_results)
            for r in itn_results:
                self.add(r)
        self.sum_neval = sum_neval

    def extend(self, ravg):
        r""" Merge results from :class:`RAvg` object ``ravg`` after results currently in ``self``. """
        for r in ravg.itn_results:
            self.add(r)
        self.sum_neval += ravg.sum_neval

    def __reduce_ex__(self, protocol):
        return (RAvg, (self.weighted, gv.dumps(self.itn_results, protocol=protocol), self.sum_neval))

    @property
    def chi2(self):
        "*chi**2* of weighted average."
        if len(self.itn_results) <= 1:
            return 0.0
        wavg = self.mean
        if self.weighted:
            ans = 0.0
            for m, w in zip(self._mlist, self._wlist):
                ans += (wavg - m) ** 2 * w
            return ans
        return np.sum([(m - wavg) ** 2 for m in self._mlist]) / (self._varsum / self._n)

    @property
    def dof(self):
        "Number of degrees of freedom in weighted average."
        return len(self.itn_results) - 1

    @property
    def nitn(self):
        "Number of iterations."
        return len(self.itn_results)

    @property
    def Q(self):
        "*Q* or *p-value* of weighted average's *chi**2*."
        return gv.gammaQ(self.dof / 2., self.chi2 / 2.) if self.dof > 0 and self.chi2 >= 0 else float('nan')

    @property
    def avg_neval(self):
        "Average number of integrand evaluations per iteration."
        return self.sum_neval / self.nitn if self.nitn > 0 else 0

    def converged(self, rtol, atol):
        return self.sdev < atol + rtol * abs(self.mean)

    def add(self, g):
        r""" Add estimate ``g`` to the running average. """
        self.itn_results.append(g)
        if isinstance(g, gv.GVarRef):
            return
        self._mlist.append(g.mean)
        if self.weighted:
            self._wlist.append(1 / (g.var if g.var > TINY else TINY))
            var = 1. / np.sum(self._wlist)
            sdev = np.sqrt(var)
            mean = np.sum([w * m for w, m in zip(self._wlist, self._mlist)]) * var
            super(RAvg, self).__init__(*gv.gvar(mean, sdev).internaldata)
        else:
            self._msum += g.mean
            self._varsum += g.var
            self._n += 1
            mean = self._msum / self._n
            var = self._varsum / self._n ** 2
            super(RAvg, self).__init__(*gv.gvar(mean, np.sqrt(var)).internaldata)

    def summary(self, extended=False, weighted=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        if weighted is None:
            weighted = self.weighted
        return _summary_table(self.itn_results, lambda: RAvg(weighted=weighted), lambda r: r, weighted)


class RAvgArray(np.ndarray):
    r""" Running average of array-valued Monte Carlo estimates (``_vegas.pyx:2579-2879``): an
    ``ndarray`` of Gaussian variables; estimates are combined with their inverse covariance
    matrices (SVD-protected) if ``weighted=True``. """

    def __new__(subtype, shape=None, dtype=object, buffer=None, offset=0, strides=None, order=None,
                weighted=True, itn_results=None, sum_neval=0, rescale=True):
        if shape is None and (itn_results is None or len(itn_results) < 1):
            raise ValueError('must specificy shape or itn_results')
        obj = np.ndarray.__new__(
            subtype, shape=shape if shape is not None else np.shape(itn_results[0]),
            dtype=object, buffer=buffer, offset=offset, strides=strides, order=order)
        if buffer is None:
            obj.flat = np.array(obj.size * [gv.gvar(0, 0)])
        obj.itn_results = []
        obj._mlist = []
        if rescale is False or rescale is None or not weighted:
            obj.rescale = None
        elif rescale is True:
            obj.rescale = True
        elif hasattr(rescale, 'keys'):
            obj.rescale = gv.asbufferdict(rescale)
        else:
            obj.rescale = np.asarray(rescale)
        if weighted:
            obj.weighted = True
            obj._wlist = []
        else:
            obj._msum = 0.
            obj._covsum = 0.
            obj._n = 0
            obj.weighted = False
        obj.sum_neval = sum_neval
        return obj

    def __reduce_ex__(self, protocol):
        save = np.array(self.flat[:])
        self.flat[:] = 0
        superpickled = super(RAvgArray, self).__reduce__()
        self.flat[:] = save
        state = superpickled[2] + (
            self.weighted, gv.dumps(self.itn_results, protocol=protocol), (self.sum_neval, self.rescale))
        return (superpickled[0], superpickled[1], state)

    def __setstate__(self, state):
        super(RAvgArray, self).__setstate__(state[:-3])
        if isinstance(state[-1], tuple):
            self.sum_neval, self.rescale = state[-1]
        else:
            self.sum_neval, self.rescale = state[-1], True
        itn_results = gv.loads(state[-2])
        self.weighted = state[-3]
        if self.weighted:
            self._wlist = []
            self._mlist = []
        else:
            self._msum = 0.
            self._covsum = 0.
            self._n = 0
            self._mlist = []
        self.__dict__.pop('_rescale', None)
        self.itn_results = []
        for r in itn_results:
            self.add(r)

    def __array_finalize__(self, obj):
        if obj is None:
            return
        if getattr(obj, 'weighted', True):
            self.weighted = True
            self._wlist = getattr(obj, '_wlist', [])
        else:
            self._msum = getattr(obj, '_msum', 0.)
            self._covsum = getattr(obj, '_covsum', 0.)
            self._n = getattr(obj, '_n', 0.)
            self.weighted = False
        self._mlist = getattr(obj, '_mlist', [])
        self.itn_results = getattr(obj, 'itn_results', [])
        self.sum_neval = getattr(obj, 'sum_neval', 0)
        self.rescale = getattr(obj, 'rescale', True)

    def __init__(self, shape=None, dtype=object, buffer=None, offset=0, strides=None, order=None,
                 weighted=True, itn_results=None, sum_neval=0, rescale=True):
        self[:] *= 0
        if itn_results is not None:
            if isinstance(itn_results, bytes):
                itn_results = gv.loads(itn_results)
            self.itn_results = []
            for r in itn_results:
                self.add(r)

    def extend(self, ravg):
        r""" Merge results from :class:`RAvgArray` object ``ravg`` after results currently in ``self``. """
        for r in ravg.itn_results:
            self.add(r)
        self.sum_neval += ravg.sum_neval

    def _w(self, matrix, rescale=False):
        " Decompose inverse matrix, with protection against singular matrices. "
        s = gv.SVD(matrix, svdcut=-EPSILON * len(matrix) * 1e4, rescale=rescale)
        return s.decomp(-1)

    def converged(self, rtol, atol):
        return np.all(gv.sdev(self) < atol + rtol * np.abs(gv.mean(self)))

    @property
    def chi2(self):
        "*chi**2* of weighted average."
        if len(self.itn_results) <= 1:
            return 0.0
        wavg = np.array(gv.mean(self), dtype=float).reshape((-1,))
        ans = 0.0
        if self.weighted:
            if self.rescale is not None:
                wavg = wavg / self._rescale
            for w, m in zip(self._wlist, self._mlist):
                for wi in w:
                    ans += wi.dot(m - wavg) ** 2
            return ans
        if self._invw is None:
            self._invw = self._w(self._covsum / self._n)
        for m in self._mlist:
            delta = wavg - m
            for invwi in self._invw:
                ans += invwi.dot(delta) ** 2
        return ans

    @property
    def dof(self):
        "Number of degrees of freedom in weighted average."
        if len(self.itn_results) <= 1:
            return 0
        if not self.weighted:
            if self._invw is None:
                self._invw = self._w(self._covsum / self._n)
            return (len(self.itn_results) - 1) * len(self._invw)
        return np.sum([len(w) for w in self._wlist]) - self.size

    @property
    def nitn(self):
        "Number of iterations."
        return len(self.itn_results)

    @property
    def Q(self):
        "*Q* or *p-value* of weighted average's *chi**2*."
        if self.dof <= 0 or self.chi2 < 0:
            return float('nan')
        return gv.gammaQ(self.dof / 2., self.chi2 / 2.)

    @property
    def avg_neval(self):
        "Average number of integrand evaluations per iteration."
        return self.sum_neval / self.nitn if self.nitn > 0 else 0

    def add(self, g):
        r""" Add estimate ``g`` to the running average. """
        g = np.asarray(g)
        self.itn_results.append(g)
        if g.size > 1 and isinstance(g.flat[0], gv.GVarRef):
            return
        g = g.reshape((-1,))
        if self.weighted:
            if '_rescale' not in self.__dict__:
                if self.rescale is not None:
                    self._rescale = np.fabs(gv.mean(g if self.rescale is True else self.rescale.flat[:]))
                    gsdev = gv.sdev(g)
                    idx = gsdev > self._rescale
                    self._rescale[idx] = gsdev[idx]
                    self._rescale[self._rescale <= 0] = 1.
                else:
                    self._rescale = 1.
            g = g / self._rescale
            gmean = gv.mean(g)
            gcov = gv.evalcov(g)
            for i in range(len(gcov)):
                if gcov[i, i] <= 0:
                    gcov[i, i] = TINY
            self._mlist.append(gmean)
            self._wlist.append(self._w(gcov))
            invcov = np.sum([(w.T).dot(w) for w in self._wlist], axis=0)
            invw = self._w(invcov)
            cov = (invw.T).dot(invw)
            mean = 0.0
            for m, w in zip(self._mlist, self._wlist):
                for wj in w:
                    wj_m = wj.dot(m)
                    for invwi in invw:
                        mean += invwi * invwi.dot(wj) * wj_m
            self[:] = (gv.gvar(mean, cov) * self._rescale).reshape(self.shape)
        else:
            gmean = gv.mean(g)
            gcov = gv.evalcov(g)
            self._mlist.append(gmean)
            self._msum += gmean
            self._covsum += gcov
            self._invw = None
            self._n += 1
            mean = self._msum / self._n
            cov = self._covsum / (self._n ** 2)
            self[:] = gv.gvar(mean, cov).reshape(self.shape)

    def summary(self, extended=False, weighted=None, rescale=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        if weighted is None:
            weighted = self.weighted
        if rescale is None:
            rescale = self.rescale
        ans = _summary_table(self.itn_results,
                             lambda: RAvgArray(self.shape, weighted=weighted, rescale=rescale),
                             lambda r: r.flat[0], weighted)
        if extended and self.itn_results[0].size > 1:
            ans += '\n' + gv.tabulate(self) + '\n'
        return ans


class RAvgDict(gv.BufferDict):
    r""" Running average of dictionary-valued Monte Carlo estimates (``_vegas.pyx:2453-2577``); the
    values share one flat :class:`RAvgArray`. """

    def __init__(self, dictionary=None, weighted=True, itn_results=None, sum_neval=0, rescale=True):
        if isinstance(itn_results, bytes):
            itn_results = gv.loads(itn_results)
        if dictionary is None and (itn_results is None or len(itn_results) < 1):
            raise ValueError('must specificy dictionary or itn_results')
        super(RAvgDict, self).__init__(dictionary if dictionary is not None else itn_results[0])
        self.rarray = RAvgArray(shape=(self.size,), weighted=weighted, rescale=rescale)
        self.buf = np.asarray(self.rarray)
        self.itn_results = []
        self.weighted = weighted
        if itn_results is not None:
            for r in itn_results:
                self.add(r)
        self.sum_neval = sum_neval

    def extend(self, ravg):
        r""" Merge results from :class:`RAvgDict` object ``ravg`` after results currently in ``self``. """
        for r in ravg.itn_results:
            self.add(r)
        self.sum_neval += ravg.sum_neval

    def __reduce_ex__(self, protocol):
        return (RAvgDict, (None, self.weighted, gv.dumps(self.itn_results, protocol=protocol),
                           self.sum_neval, self.rescale))

    def converged(self, rtol, atol):
        return np.all(gv.sdev(self.buf) < atol + rtol * np.abs(gv.mean(self.buf)))

    def add(self, g):
        if isinstance(g, gv.BufferDict):
            newg = gv.BufferDict(g)
        else:
            newg = gv.BufferDict()
            for k in self:
                try:
                    newg[k] = g[k]
                except (AttributeError, KeyError):
                    raise ValueError("Dictionary g doesn't contain key " + str(k) + '.')
        self.itn_results.append(newg)
        self.rarray.add(newg.buf)

    def summary(self, extended=False, weighted=None, rescale=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        if weighted is None:
            weighted = self.weighted
        if rescale is None:
            rescale = self.rarray.rescale
        ans = self.rarray.summary(weighted=weighted, extended=False, rescale=rescale)
        if extended and self.itn_results[0].size > 1:
            ans += '\n' + gv.tabulate(self) + '\n'
        return ans

    chi2 = property(lambda self: self.rarray.chi2, None, None, "*chi**2* of weighted average.")
    dof = property(lambda self: self.rarray.dof, None, None, "Number of degrees of freedom in weighted average.")
    nitn = property(lambda self: len(self.itn_results), None, None, "Number of iterations.")
    Q = property(lambda self: self.rarray.Q, None, None, "*Q* or *p-value* of weighted average's *chi**2*.")
    avg_neval = property(lambda self: self.sum_neval / self.nitn if self.nitn > 0 else 0, None, None,
                         "Average number of integrand evaluations per iteration.")
    rescale = property(lambda self: self.rarray.rescale, None, None,
                       "Integrals divided by ``rescale`` before doing weighted averages.")


class VegasResult(object):
    """ Accumulated result object --- standard interface for integration results
    (``_vegas.pyx:2892-2957``): picks RAvg / RAvgArray / RAvgDict from the integrand's shape. """

    def __init__(self, integrand=None, weighted=None):
        self.integrand = integrand
        self.shape = integrand.shape
        self.sum_neval = 0
        if self.shape is None:
            self.result = RAvgDict(integrand.bdict, weighted=weighted)
        elif self.shape == ():
            self.result = RAvg(weighted=weighted)
        else:
            self.result = RAvgArray(self.shape, weighted=weighted)

    def save(self, outfile):
        " pickle current results in ``outfile`` for later use. "
        if isinstance(outfile, str):
            with open(outfile, 'wb') as ofile:
                pickle.dump(self.result, ofile)
        else:
            pickle.dump(self.result, outfile)

    def saveall(self, integrator, outfile):
        " pickle current (results,integrator) in ``outfile`` for later use. "
        if isinstance(outfile, str):
            with open(outfile, 'wb') as ofile:
                pickle.dump((self.result, integrator), ofile)
        else:
            pickle.dump((self.result, integrator), outfile)

    def update(self, mean, var, last_neval=None):
        self.result.add(self.integrand.format_result(mean, var))
        if last_neval is not None:
            self.sum_neval += last_neval
            self.result.sum_neval = self.sum_neval

    def update_analyzer(self, analyzer):
        r""" Update analyzer at end of an iteration. """
        analyzer.end(self.result.itn_results[-1], self.result)

    def converged(self, rtol, atol):
        " Convergence test. "
        return self.result.converged(rtol, atol)


class reporter(object):
    r""" Analyzer class that prints out a report, iteration by iteration, on how vegas is doing
    (``_vegas.pyx:2232-2269``).  ``ngrid`` = number of grid nodes printed per direction. """

    def __init__(self, ngrid=0):
        self.ngrid = ngrid
        self.clock = time.perf_counter

    def begin(self, itn, integrator):
        self.integrator = integrator
        self.itn = itn
        self.t0 = self.clock()
        if itn == 0:
            print(integrator.settings())
        sys.stdout.flush()

    def end(self, itn_ans, ans):
        print("    itn %2d: %s\n all itn's: %s" % (self.itn + 1, itn_ans, ans))
        print('    neval = %s  neval/h-cube = %s\n    chi2/dof = %.2f  Q = %.2f  time = %.2f' % (
            format(self.integrator.last_neval, '.6g'),
            tuple(self.integrator.neval_hcube_range),
            ans.chi2 / ans.dof if ans.dof > 0 else 0,
            ans.Q if ans.dof > 0 else 1.,
            self.clock() - self.t0))
        print(self.integrator.map.settings(ngrid=self.ngrid))
        print('')
        sys.stdout.flush()


def ravg(reslist, weighted=None, rescale=None):
    r""" Running average built from a list of |vegas| results (``src/vegas/__init__.py:1220-1311``):
    ``reslist`` is a list of |GVar|\s, arrays of them, or dictionaries -- or a result object, whose
    ``itn_results`` are then used.  ``weighted=False`` gives the unweighted (unbiased) average, e.g.
    ``vegas.ravg(r.itn_results[5:], weighted=False)`` to drop the iterations where the map was still
    adapting.  ``rescale`` as in :class:`RAvgArray`. """
    from . import _pdf                     # (imported here: _pdf builds on this module)
    for t in (_pdf.PDFEV, _pdf.PDFEVArray, _pdf.PDFEVDict):
        if isinstance(reslist, t):         # average the underlying integrals, then form the ratios again
            return t(ravg(reslist.itn_results, weighted=weighted, rescale=rescale))
    src = reslist
    if isinstance(reslist, (RAvg, RAvgArray, RAvgDict)):
        reslist = reslist.itn_results
    try:
        if len(reslist) < 1:
            raise ValueError('reslist empty')
    except TypeError:
        raise ValueError('improper type for reslist')
    if weighted is None:
        weighted = getattr(src, 'weighted', True)
    if rescale is None:
        rescale = getattr(src, 'rescale', reslist[-1])
    if hasattr(reslist[0], 'keys'):
        return RAvgDict(itn_results=reslist, weighted=weighted, rescale=rescale)
    try:
        shape = np.shape(reslist[0])
    except Exception:
        raise ValueError('reslist[i] not GVar, array, or dictionary')
    if shape == ():
        return RAvg(itn_results=reslist, weighted=weighted)
    return RAvgArray(itn_results=reslist, weighted=weighted, rescale=rescale)
