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


# --------------------------------------------------------------------------- the averaging core
class _ScalarAverage(object):
    """running average of scalar estimates (mean_i, var_i): inverse-variance weighted, or plain
    (formulas of ``_vegas.pyx:2392-2410``).  O(1) per estimate: running sums, not lists re-summed."""

    def __init__(self, weighted):
        self.weighted = bool(weighted)
        self.means, self.weights = [], []          # kept for chi2 (computed on demand)
        self.sw = self.swm = 0.0                   # weighted: sum w, sum w m
        self.msum = self.varsum = 0.0              # unweighted
        self.n = 0

    def add(self, mean, var):
        """returns (average, its variance) after taking in one estimate"""
        self.n += 1
        self.means.append(mean)
        if self.weighted:
            w = 1. / (var if var > TINY else TINY)
            self.weights.append(w)
            self.sw += w
            self.swm += w * mean
            return self.swm / self.sw, 1. / self.sw
        self.msum += mean
        self.varsum += var
        return self.msum / self.n, self.varsum / self.n ** 2

    def chi2(self, avg):
        if self.n <= 1:
            return 0.0
        if self.weighted:
            return float(sum((avg - m) ** 2 * w for m, w in zip(self.means, self.weights)))
        return float(sum((m - avg) ** 2 for m in self.means) / (self.varsum / self.n))

    @property
    def dof(self):
        return self.n - 1


class _VectorAverage(object):
    """running average of vector estimates (mean_i[n], cov_i[n, n]) (formulas of ``_vegas.pyx:2801-2846``).

    Weighted: every estimate contributes ``W_i = svd-protected decomposition of cov_i^-1`` (rows ``w`` with
    ``cov_i^-1 = sum_w w w^T`` over the modes kept); the average is ``C sum_i W_i^T W_i m_i`` with
    ``C = (sum_i W_i^T W_i)^-1``, again SVD-protected.  The two sums are kept running.  Estimates are divided
    by ``scale`` first (``rescale``: fixed at the first estimate), which keeps the matrices well conditioned
    when components differ by orders of magnitude."""

    def __init__(self, weighted, rescale):
        self.weighted = bool(weighted)
        self.rescale = rescale                    # None | True | array of typical values
        self.scale = None                         # set at the first estimate (weighted only)
        self.means, self.ws = [], []
        self.S = self.b = None                    # weighted: sum W^T W, sum W^T W m
        self.msum = self.covsum = 0.0             # unweighted
        self.n = 0
        self._invw = None

    @staticmethod
    def decomp(matrix, rescale=False):
        " rows w with matrix^-1 = sum w w^T, modes with tiny eigenvalues dropped "
        return gv.SVD(matrix, svdcut=-EPSILON * len(matrix) * 1e4, rescale=rescale).decomp(-1)

    def _set_scale(self, mean, sdev):
        if self.rescale is None:
            self.scale = 1.
            return
        sc = np.fabs(mean if self.rescale is True else np.asarray(gv.mean(self.rescale.flat[:]), float))
        sc = np.array(sc, dtype=float)
        big = sdev > sc
        sc[big] = sdev[big]
        sc[sc <= 0] = 1.
        self.scale = sc

    def add(self, mean, cov):
        """returns (average[n], covariance[n, n]) after taking in one estimate"""
        self.n += 1
        if not self.weighted:
            self.means.append(mean)
            self.msum = self.msum + mean
            self.covsum = self.covsum + cov
            self._invw = None
            return self.msum / self.n, self.covsum / self.n ** 2
        if self.scale is None:
            self._set_scale(mean, np.sqrt(np.fabs(np.diag(cov))))
        mean = mean / self.scale
        cov = cov / np.outer(self.scale, self.scale) if np.ndim(self.scale) else np.array(cov, float)
        d = np.diag_indices(len(cov))
        cov[d] = np.where(cov[d] <= 0, TINY, cov[d])
        w = self.decomp(cov)
        self.means.append(mean)
        self.ws.append(w)
        wtw = w.T.dot(w)
        self.S = wtw if self.S is None else self.S + wtw
        wm = wtw.dot(mean)
        self.b = wm if self.b is None else self.b + wm
        invw = self.decomp(self.S)
        avg_cov = invw.T.dot(invw)
        avg = avg_cov.dot(self.b)
        sc = self.scale
        return avg * sc, avg_cov * (np.outer(sc, sc) if np.ndim(sc) else 1.)

    def chi2(self, avg):
        if self.n <= 1:
            return 0.0
        ans = 0.0
        if self.weighted:
            avg = avg / self.scale
            for w, m in zip(self.ws, self.means):
                ans += float(np.sum(w.dot(m - avg) ** 2))
            return ans
        if self._invw is None:
            self._invw = self.decomp(self.covsum / self.n)
        for m in self.means:
            ans += float(np.sum(self._invw.dot(avg - m) ** 2))
        return ans

    def dof(self, size):
        if self.n <= 1:
            return 0
        if not self.weighted:
            if self._invw is None:
                self._invw = self.decomp(self.covsum / self.n)
            return (self.n - 1) * len(self._invw)
        return int(sum(len(w) for w in self.ws)) - size


def _rescale_arg(rescale, weighted):
    if rescale is False or rescale is None or not weighted:
        return None
    if rescale is True:
        return True
    return gv.asbufferdict(rescale) if hasattr(rescale, 'keys') else np.asarray(rescale)


class _RunningAverage(object):
    """what RAvg, RAvgArray and RAvgDict share: the list of per-iteration results, the neval count, and
    the statistics of the average (chi2, dof, Q) from the core in ``self._avg``"""

    def extend(self, ravg):
        r""" Merge results from ``ravg`` after results currently in ``self``. """
        for r in ravg.itn_results:
            self.add(r)
        self.sum_neval += ravg.sum_neval

    @property
    def nitn(self):
        "Number of iterations."
        return len(self.itn_results)

    @property
    def avg_neval(self):
        "Average number of integrand evaluations per iteration."
        return self.sum_neval / self.nitn if self.nitn > 0 else 0

    @property
    def Q(self):
        "*Q* or *p-value* of weighted average's *chi**2*."
        dof, chi2 = self.dof, self.chi2
        return gv.gammaQ(dof / 2., chi2 / 2.) if dof > 0 and chi2 >= 0 else float('nan')

    def _table(self, fresh, first, weighted, extended):
        """iteration-by-iteration table: each result next to the running average up to it"""
        acc, rows = fresh(), []
        for i, res in enumerate(self.itn_results):
            acc.add(res)
            rows.append(('%3d' % (i + 1), '%-15s' % first(res), '%-15s' % first(acc),
                         '%8.2f' % (acc.chi2 / acc.dof if i != 0 else 0.0), '%8.2f' % (acc.Q if i != 0 else 1.0)))
        width = [max(len(r[c]) for r in rows) if rows else 0 for c in range(5)]
        fmt = '%%%ds   %%-%ds %%-%ds %%%ds %%%ds\n' % tuple(width)
        out = fmt % ('itn', 'integral', 'wgt average' if weighted else 'average', 'chi2/dof', 'Q')
        out += len(out[:-1]) * '-' + '\n'
        for r in rows:
            out += fmt % r
        if extended and np.size(self.itn_results[0]) > 1:
            out += '\n' + gv.tabulate(self) + '\n'
        return out


class RAvg(_RunningAverage, gv.GVar):
    r""" Running average of scalar-valued Monte Carlo estimates (API of ``_vegas.pyx:2276-2451``).

    Estimates are weighted by their inverse variances if ``weighted=True``; otherwise straight,
    unweighted averages are used. """

    def __init__(self, weighted=True, itn_results=None, sum_neval=0, _rescale=True):
        self.rescale = None
        self.weighted = bool(weighted)
        self._avg = _ScalarAverage(weighted)
        self.itn_results = []
        gv.GVar.__init__(self, *gv.gvar(0., 0.).internaldata)
        if itn_results is not None:
            for r in (gv.loads(itn_results) if isinstance(itn_results, bytes) else itn_results):
                self.add(r)
        self.sum_neval = sum_neval

    def __reduce_ex__(self, protocol):
        return (RAvg, (self.weighted, gv.dumps(self.itn_results, protocol=protocol), self.sum_neval))

    def add(self, g):
        r""" Add estimate ``g`` to the running average. """
        self.itn_results.append(g)
        if isinstance(g, gv.GVarRef):
            return
        mean, var = self._avg.add(g.mean, g.var)
        gv.GVar.__init__(self, *gv.gvar(mean, np.sqrt(var)).internaldata)

    chi2 = property(lambda self: self._avg.chi2(self.mean), None, None, "*chi**2* of weighted average.")
    dof = property(lambda self: len(self.itn_results) - 1, None, None, "Number of degrees of freedom in weighted average.")

    def converged(self, rtol, atol):
        return self.sdev < atol + rtol * abs(self.mean)

    def summary(self, extended=False, weighted=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        weighted = self.weighted if weighted is None else weighted
        return self._table(lambda: RAvg(weighted=weighted), lambda r: r, weighted, False)


def _rebuild_array(shape, weighted, itn_results, sum_neval, rescale):
    return RAvgArray(shape, weighted=weighted, itn_results=itn_results, sum_neval=sum_neval, rescale=rescale)


class RAvgArray(_RunningAverage, np.ndarray):
    r""" Running average of array-valued Monte Carlo estimates (API of ``_vegas.pyx:2579-2879``): an
    ``ndarray`` of Gaussian variables; estimates are combined with their inverse covariance matrices
    (SVD-protected) if ``weighted=True``.  ``rescale``: integrals are divided by ``rescale`` (``True``: by
    the first estimate) before weighted averages are taken. """

    def __new__(subtype, shape=None, dtype=object, buffer=None, offset=0, strides=None, order=None,
                weighted=True, itn_results=None, sum_neval=0, rescale=True):
        if isinstance(itn_results, bytes):
            itn_results = gv.loads(itn_results)
        if shape is None and (itn_results is None or len(itn_results) < 1):
            raise ValueError('must specificy shape or itn_results')
        obj = np.ndarray.__new__(
            subtype, shape=shape if shape is not None else np.shape(itn_results[0]),
            dtype=object, buffer=buffer, offset=offset, strides=strides, order=order)
        if buffer is None:
            obj.flat = np.array(obj.size * [gv.gvar(0, 0)])
        obj.weighted = bool(weighted)
        obj.rescale = _rescale_arg(rescale, weighted)
        obj._avg = _VectorAverage(weighted, obj.rescale)
        obj.itn_results = []
        obj.sum_neval = sum_neval
        return obj

    def __init__(self, shape=None, dtype=object, buffer=None, offset=0, strides=None, order=None,
                 weighted=True, itn_results=None, sum_neval=0, rescale=True):
        if itn_results is not None:
            for r in (gv.loads(itn_results) if isinstance(itn_results, bytes) else itn_results):
                self.add(r)

    def __array_finalize__(self, obj):
        if obj is None:
            return
        # views and copies share the bookkeeping of the array they come from
        self.weighted = getattr(obj, 'weighted', True)
        self.rescale = getattr(obj, 'rescale', True)
        self._avg = getattr(obj, '_avg', None)
        self.itn_results = getattr(obj, 'itn_results', [])
        self.sum_neval = getattr(obj, 'sum_neval', 0)

    def __reduce_ex__(self, protocol):
        return (_rebuild_array, (self.shape, self.weighted, gv.dumps(self.itn_results, protocol=protocol),
                                 self.sum_neval, self.rescale))

    def add(self, g):
        r""" Add estimate ``g`` to the running average. """
        g = np.asarray(g)
        self.itn_results.append(g)
        if g.size > 1 and isinstance(g.flat[0], gv.GVarRef):
            return
        g = g.reshape(-1)
        if self._avg is None:
            self._avg = _VectorAverage(self.weighted, self.rescale)
        mean, cov = self._avg.add(np.asarray(gv.mean(g), float), np.array(gv.evalcov(g), float))
        self[...] = gv.gvar(mean, cov).reshape(self.shape)

    @property
    def chi2(self):
        "*chi**2* of weighted average."
        return self._avg.chi2(np.array(gv.mean(self), dtype=float).reshape(-1)) if len(self.itn_results) > 1 else 0.0

    @property
    def dof(self):
        "Number of degrees of freedom in weighted average."
        return self._avg.dof(self.size) if len(self.itn_results) > 1 else 0

    def converged(self, rtol, atol):
        return np.all(gv.sdev(self) < atol + rtol * np.abs(gv.mean(self)))

    def summary(self, extended=False, weighted=None, rescale=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        weighted = self.weighted if weighted is None else weighted
        rescale = self.rescale if rescale is None else rescale
        return self._table(lambda: RAvgArray(self.shape, weighted=weighted, rescale=rescale),
                           lambda r: r.flat[0], weighted, extended)


class RAvgDict(_RunningAverage, gv.BufferDict):
    r""" Running average of dictionary-valued Monte Carlo estimates (API of ``_vegas.pyx:2453-2577``); the
    values share one flat :class:`RAvgArray` (``rarray``). """

    def __init__(self, dictionary=None, weighted=True, itn_results=None, sum_neval=0, rescale=True):
        if isinstance(itn_results, bytes):
            itn_results = gv.loads(itn_results)
        if dictionary is None and (itn_results is None or len(itn_results) < 1):
            raise ValueError('must specificy dictionary or itn_results')
        gv.BufferDict.__init__(self, dictionary if dictionary is not None else itn_results[0])
        self.rarray = RAvgArray(shape=(self.size,), weighted=weighted, rescale=rescale)
        self.buf = np.asarray(self.rarray)
        self.itn_results = []
        self.weighted = weighted
        for r in (itn_results or []):
            self.add(r)
        self.sum_neval = sum_neval

    def __reduce_ex__(self, protocol):
        return (RAvgDict, (None, self.weighted, gv.dumps(self.itn_results, protocol=protocol),
                           self.sum_neval, self.rescale))

    def add(self, g):
        r""" Add estimate ``g`` (a dictionary with this dictionary's keys) to the running average. """
        if not isinstance(g, gv.BufferDict):
            missing = [k for k in self if k not in g]
            if missing or not hasattr(g, 'keys'):
                raise ValueError("Dictionary g doesn't contain key " + str(missing[0] if missing else None) + '.')
            g = {k: g[k] for k in self}                 # this dictionary's key order
        entry = gv.BufferDict(g)
        self.itn_results.append(entry)
        self.rarray.add(entry.buf)

    chi2 = property(lambda self: self.rarray.chi2, None, None, "*chi**2* of weighted average.")
    dof = property(lambda self: self.rarray.dof, None, None, "Number of degrees of freedom in weighted average.")
    rescale = property(lambda self: self.rarray.rescale, None, None,
                       "Integrals divided by ``rescale`` before doing weighted averages.")

    def converged(self, rtol, atol):
        return np.all(gv.sdev(self.buf) < atol + rtol * np.abs(gv.mean(self.buf)))

    def summary(self, extended=False, weighted=None, rescale=None):
        r""" Assemble summary of results, iteration-by-iteration, into a string. """
        weighted = self.weighted if weighted is None else weighted
        rescale = self.rarray.rescale if rescale is None else rescale
        ans = self.rarray.summary(weighted=weighted, extended=False, rescale=rescale)
        if extended and self.itn_results[0].size > 1:
            ans += '\n' + gv.tabulate(self) + '\n'
        return ans


def _dump(payload, outfile):
    if isinstance(outfile, str):
        with open(outfile, 'wb') as ofile:
            ofile.write(payload)
    else:
        outfile.write(payload)


class VegasResult(object):
    """ Accumulated result of an integration (API of ``_vegas.pyx:2892-2957``): the running average that matches
    the integrand's output -- ``RAvgDict`` for dictionaries (``shape is None``), ``RAvg`` for scalars, ``RAvgArray``
    otherwise -- plus the running count of integrand evaluations. """

    def __init__(self, integrand=None, weighted=None):
        self.integrand, self.shape, self.sum_neval = integrand, integrand.shape, 0
        make = {None: lambda: RAvgDict(integrand.bdict, weighted=weighted), (): lambda: RAvg(weighted=weighted)}
        self.result = make.get(self.shape, lambda: RAvgArray(self.shape, weighted=weighted))()

    # With the hypercube range sharded over ranks every rank holds the same results; only the writer
    # (rank 0, as in the reference: pyx:2923-2941 with mpi_rank) touches the file.  Pickling the integrator
    # gathers sigf from all ranks -- a collective -- so it happens on every rank before the writer writes.
    is_writer = True

    def save(self, outfile):
        " pickle current results in ``outfile`` for later use. "
        if self.is_writer:
            _dump(pickle.dumps(self.result), outfile)

    def saveall(self, integrator, outfile):
        " pickle current (results,integrator) in ``outfile`` for later use. "
        payload = pickle.dumps((self.result, integrator))
        if self.is_writer:
            _dump(payload, outfile)

    def update(self, mean, var, last_neval=None):
        """fold one iteration's estimate (flat ``mean``, covariance or variances ``var``) into the average"""
        self.result.add(self.integrand.format_result(mean, var))
        if last_neval is not None:
            self.sum_neval += last_neval
            self.result.sum_neval = self.sum_neval

    def update_analyzer(self, analyzer):
        r""" Hand the latest iteration and the running average to ``analyzer.end``. """
        analyzer.end(self.result.itn_results[-1], self.result)

    def converged(self, rtol, atol):
        " Has the average reached the requested tolerances? "
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
        I = self.integrator
        have_dof = ans.dof > 0
        print("    itn %2d: %s\n all itn's: %s" % (self.itn + 1, itn_ans, ans))
        print('    neval = %s  neval/h-cube = %s\n    chi2/dof = %.2f  Q = %.2f  time = %.2f' % (
            format(I.last_neval, '.6g'), tuple(I.neval_hcube_range), ans.chi2 / ans.dof if have_dof else 0,
            ans.Q if have_dof else 1., self.clock() - self.t0))
        print(I.map.settings(ngrid=self.ngrid))
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
    src, items = reslist, reslist
    if isinstance(reslist, (RAvg, RAvgArray, RAvgDict)):
        items = reslist.itn_results
    try:
        n = len(items)
    except TypeError:
        raise ValueError('improper type for reslist')
    if n < 1:
        raise ValueError('reslist empty')
    weighted = getattr(src, 'weighted', True) if weighted is None else weighted
    rescale = getattr(src, 'rescale', items[-1]) if rescale is None else rescale
    first = items[0]
    if hasattr(first, 'keys'):
        return RAvgDict(itn_results=items, weighted=weighted, rescale=rescale)
    try:
        scalar = np.shape(first) == ()
    except Exception:
        raise ValueError('reslist[i] not GVar, array, or dictionary')
    if scalar:
        return RAvg(itn_results=items, weighted=weighted)
    return RAvgArray(itn_results=items, weighted=weighted, rescale=rescale)
