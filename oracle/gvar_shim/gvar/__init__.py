"""Minimal stand-in for the third-party ``gvar`` package (pin: gvar>=13.1.5,
/root/reference/pyproject.toml:28), which is not installed in this image.

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference
extension built into ``oracle/_ref`` can be imported (``_vegas.pyx:32`` does
``import gvar``) and driven as a CPU oracle / CPU baseline.  Product code under
``vegas_b200/`` never imports this module.

Only the behaviour the reference's hot path touches is provided:

* ``RNG`` / ``ranseed``            -- numpy Generator (``_vegas.pyx:1215,1677``)
* ``GVar``, ``gvar``, ``mean``, ``sdev``, ``var``, ``evalcov``, ``evalcorr``
                                    -- Gaussian variables with *linear* correlation
                                       tracking through ``+ - * /`` by constants
                                       (``_vegas.pyx:2403,2801-2818,3143-3147``)
* ``BufferDict`` / ``asbufferdict`` -- ordered dict over one flat buffer
                                       (``_vegas.pyx:1225-1237,3132,3158``)
* ``SVD``                           -- eigen-decomposition with svdcut
                                       (``_vegas.pyx:2721``)
* ``gammaQ``, ``GVarRef``, ``dumps`` / ``loads`` (pickle based)
* ``PDF``, ``PDFStatistics``, ``gvar('1.0(5)')``  -- what ``vegas.PDFIntegrator`` needs
                                       (``src/vegas/__init__.py:496-503, 599-606, 1009``)

The real gvar algorithms are re-stated from their published behaviour; nothing
here is used to produce shipped results.
"""
import pickle as _pickle
import re as _re

import numpy as _np
from scipy.special import gammaincc as _gammaincc

__version__ = '0.0-shim'

RNG = _np.random.default_rng()


def ranseed(seed=None, size=3, version=None):
    """Seed the shared generator; returns the seed used (tuple)."""
    global RNG
    if seed is None:
        seed = tuple(int(s) for s in _np.random.SeedSequence().generate_state(size))
    try:
        seed = tuple(int(s) for s in seed)
    except TypeError:
        seed = (int(seed),)
    RNG = _np.random.default_rng(list(seed))
    # keep the legacy numpy global stream seeded as well (reference tests use it)
    _np.random.seed(int(sum(seed)) % (2 ** 32))
    return seed


class GVarRef(object):
    """Placeholder type: the reference only does isinstance checks on it."""


class _Block(object):
    """Shared covariance block for GVars created together."""
    __slots__ = ('cov',)

    def __init__(self, cov):
        self.cov = _np.array(cov, dtype=float)


class GVar(object):
    """Gaussian variable = mean + sum_i coef_i * z_i with z ~ N(0, block.cov)."""

    def __init__(self, mean=0.0, terms=None, *extra):
        # ``internaldata`` round-trips through this signature (RAvg.__init__)
        self.mean = float(mean)
        self._terms = list(terms) if terms else []   # [(block, index, coef)]

    @property
    def internaldata(self):
        return (self.mean, self._terms)

    @property
    def var(self):
        return float(_cov_between(self, self))

    @property
    def sdev(self):
        v = self.var
        return float(v ** 0.5) if v > 0 else 0.0

    # --- linear arithmetic -------------------------------------------------
    def _scaled(self, c):
        return GVar(self.mean * c, [(b, i, k * c) for b, i, k in self._terms])

    def __neg__(self):
        return self._scaled(-1.0)

    def __pos__(self):
        return self

    def __add__(self, other):
        if isinstance(other, GVar):
            return GVar(self.mean + other.mean, self._terms + other._terms)
        if isinstance(other, _np.ndarray):
            return NotImplemented
        return GVar(self.mean + float(other), self._terms)
    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, _np.ndarray):
            return NotImplemented
        return self + (-other)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if isinstance(other, GVar):
            a = self._scaled(other.mean)
            b = other._scaled(self.mean)
            return GVar(self.mean * other.mean, a._terms + b._terms)
        if isinstance(other, _np.ndarray):
            return NotImplemented
        return self._scaled(float(other))
    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, GVar):
            inv = GVar(1.0 / other.mean,
                       [(b, i, -k / other.mean ** 2) for b, i, k in other._terms])
            return self * inv
        if isinstance(other, _np.ndarray):
            return NotImplemented
        return self._scaled(1.0 / float(other))

    def __rtruediv__(self, other):
        inv = GVar(1.0 / self.mean,
                   [(b, i, -k / self.mean ** 2) for b, i, k in self._terms])
        return inv * other

    def __pow__(self, p):
        p = float(p)
        return GVar(self.mean ** p,
                    [(b, i, k * p * self.mean ** (p - 1)) for b, i, k in self._terms])

    # comparisons act on the means (as in gvar)
    def __lt__(self, other):
        return self.mean < getattr(other, 'mean', other)

    def __gt__(self, other):
        return self.mean > getattr(other, 'mean', other)

    def __le__(self, other):
        return self.mean <= getattr(other, 'mean', other)

    def __ge__(self, other):
        return self.mean >= getattr(other, 'mean', other)

    def __abs__(self):
        return self if self.mean >= 0 else -self

    def __call__(self):
        """A random sample from the distribution."""
        return self.mean + self.sdev * RNG.standard_normal()

    def __float__(self):
        return self.mean

    def __str__(self):
        return fmt_gvar(self.mean, self.sdev)

    __repr__ = __str__

    def __format__(self, spec):
        return format(str(self), spec)


def _cov_between(a, b):
    tot = 0.0
    for ba, ia, ka in a._terms:
        for bb, ib, kb in b._terms:
            if ba is bb:
                tot += ka * kb * ba.cov[ia, ib]
    return tot


def fmt_gvar(mean, sdev):
    """gvar's compact '1.35(86)' notation (2 significant digits on the error)."""
    if not _np.isfinite(mean) or not _np.isfinite(sdev):
        return '%g +- %g' % (mean, sdev)
    if sdev == 0:
        return '%g(0)' % mean
    if sdev < 0:
        sdev = -sdev
    # exponent of the leading digit of sdev, keep 2 digits
    e = int(_np.floor(_np.log10(sdev)))
    s2 = round(sdev / 10.0 ** (e - 1))
    if s2 >= 100:
        e += 1
        s2 = round(sdev / 10.0 ** (e - 1))
    ndec = -(e - 1)                      # decimals needed to show 2 digits
    if abs(mean) >= 1e16 * sdev or abs(mean) >= 1e7 or (0 < abs(mean) < 1e-5 and sdev < 1e-5):
        # scientific fallback
        me = int(_np.floor(_np.log10(max(abs(mean), sdev))))
        return fmt_gvar(mean / 10.0 ** me, sdev / 10.0 ** me) + 'e%+03d' % me
    if ndec <= 0:
        # error >= 10: integers
        return '%.0f(%.0f)' % (mean, round(sdev))
    if e >= 0:
        # 1 <= sdev < 10  ->  '13.5(8.6)'
        return '%.*f(%.*f)' % (ndec, mean, ndec, sdev)
    return '%.*f(%d)' % (ndec, mean, int(s2))


_GV_STR = _re.compile(r'^\s*([-+]?)(\d*)(?:\.(\d*))?\s*\(\s*(\d*)(?:\.(\d*))?\s*\)\s*(?:[eE]([-+]?\d+))?\s*$')
_GV_PM = _re.compile(r'^\s*(\S+)\s*(?:\+-|\+/-|\u00b1)\s*(\S+)\s*$')


def _parse(text):
    """'1.35(86)', '13.5(8.6)', '1(1)e-3', '1.2 +- 0.3' -> (mean, sdev)"""
    m = _GV_PM.match(text)
    if m:
        return float(m.group(1)), float(m.group(2))
    m = _GV_STR.match(text)
    if not m:
        raise ValueError('cannot convert %r to a GVar' % (text,))
    sign, ip, fp, eip, efp, ex = m.groups()
    mean_ = float((ip or '0') + '.' + (fp or '0'))
    if efp is not None:                       # error written with its own decimal point: literal
        sdev_ = float((eip or '0') + '.' + (efp or '0'))
    else:                                     # error in units of the last digit of the mean
        sdev_ = float(eip or '0') * 10.0 ** (-len(fp or ''))
    scale = 10.0 ** int(ex) if ex else 1.0
    if sign == '-':
        mean_ = -mean_
    return mean_ * scale, sdev_ * scale


def gvar(*args):
    """gvar(mean, sdev) | gvar(mean_array, sdev_array | cov_matrix) | gvar('1.0(5)') | gvar(array or dict of
    those) | gvar(GVar)."""
    if len(args) == 1:
        a = args[0]
        if isinstance(a, GVar):
            return a
        if isinstance(a, str):
            return gvar(*_parse(a))
        if hasattr(a, 'keys'):
            out = BufferDict()
            for k in a:
                out[k] = gvar(a[k])
            return out
        if isinstance(a, (tuple, list)) and len(a) == 2 and _np.ndim(a[0]) == 0 and not isinstance(a[0], (str, GVar)):
            return gvar(*a)
        arr = _np.asarray(a, dtype=object)
        if arr.shape == ():
            return gvar(arr.item()) if isinstance(arr.item(), str) else arr
        out = _np.empty(arr.shape, dtype=object)
        for idx in _np.ndindex(arr.shape):
            v = arr[idx]
            out[idx] = gvar(v) if isinstance(v, str) else v
        return out
    m, s = args
    if hasattr(m, 'keys'):
        m = asbufferdict(m)
        sb = asbufferdict(s).buf if hasattr(s, 'keys') else s
        return BufferDict(m, buf=gvar(_np.array(m.buf, dtype=float), _np.array(sb, dtype=float)))
    if _np.ndim(m) == 0:
        m = float(m)
        s = float(s)
        return GVar(m, [(_Block([[s * s]]), 0, 1.0)])
    m = _np.asarray(m, dtype=float)
    s = _np.asarray(s, dtype=float)
    if s.shape == m.shape:
        cov = _np.diag(s.reshape(-1) ** 2)
    else:
        cov = s.reshape(m.size, m.size)
    blk = _Block(cov)
    out = _np.empty(m.size, dtype=object)
    for i, mi in enumerate(m.reshape(-1)):
        out[i] = GVar(mi, [(blk, i, 1.0)])
    return out.reshape(m.shape)


def fabs(g):
    """elementwise absolute value (GVars keep their error)"""
    if isinstance(g, GVar):
        return abs(g)
    if _is_bd(g):
        return BufferDict(g, buf=fabs(g.buf))
    a = _np.asarray(g)
    if a.dtype != object:
        return _np.fabs(a)
    out = _np.empty(a.shape, dtype=object)
    for idx in _np.ndindex(a.shape):
        out[idx] = abs(a[idx])
    return out if out.shape != () else out.item()


def _is_bd(g):
    return isinstance(g, BufferDict)


def mean(g):
    if isinstance(g, GVar):
        return g.mean
    if _is_bd(g):
        return BufferDict(g, buf=mean(g.buf))
    a = _np.asarray(g)
    if a.dtype != object:
        return _np.array(a, dtype=float)
    out = _np.empty(a.shape, dtype=float)
    for idx in _np.ndindex(a.shape):
        v = a[idx]
        out[idx] = v.mean if isinstance(v, GVar) else float(v)
    return out if out.shape != () else float(out)


def sdev(g):
    if isinstance(g, GVar):
        return g.sdev
    if _is_bd(g):
        return BufferDict(g, buf=sdev(g.buf))
    a = _np.asarray(g)
    out = _np.zeros(a.shape, dtype=float)
    if a.dtype == object:
        for idx in _np.ndindex(a.shape):
            v = a[idx]
            out[idx] = v.sdev if isinstance(v, GVar) else 0.0
    return out if out.shape != () else float(out)


def var(g):
    return sdev(g) ** 2


def evalcov(g):
    if _is_bd(g):
        g = g.buf
    flat = list(_np.asarray(g, dtype=object).reshape(-1))
    n = len(flat)
    cov = _np.zeros((n, n), float)
    for i in range(n):
        for j in range(i + 1):
            if isinstance(flat[i], GVar) and isinstance(flat[j], GVar):
                cov[i, j] = cov[j, i] = _cov_between(flat[i], flat[j])
    shape = _np.shape(g)
    return cov.reshape(shape + shape) if len(shape) != 1 else cov


def evalcorr(g):
    cov = evalcov(_np.asarray(g, dtype=object).reshape(-1))
    d = _np.sqrt(_np.diag(cov))
    d[d == 0] = 1.0
    return cov / d[:, None] / d[None, :]


def corr(a, b):
    c = _cov_between(a, b)
    return c / (a.sdev * b.sdev) if a.sdev > 0 and b.sdev > 0 else 0.0


def gammaQ(a, x):
    return float(_gammaincc(a, x))


class BufferDict(dict):
    """Ordered dict whose values are views into one flat buffer ``buf``.

    ``shape`` is always ``None`` (that is how the reference tells a dict from an
    array, ``_vegas.pyx:1523``).
    """
    shape = None

    def __init__(self, *args, **kargs):
        dict.__init__(self)
        self._slices = {}
        buf = kargs.pop('buf', None)
        lbatch_buf = kargs.pop('lbatch_buf', None)
        rbatch_buf = kargs.pop('rbatch_buf', None)
        src = args[0] if args else None
        if src is None and kargs:               # BufferDict(a=..., b=...)
            src, kargs = kargs, {}
        if src is None:
            self._buf = _np.zeros(0, float)
            return
        if isinstance(src, BufferDict) and (buf is not None or lbatch_buf is not None
                                            or rbatch_buf is not None):
            if lbatch_buf is not None:
                lb = _np.asarray(lbatch_buf)
                for k in src:
                    sl, sh = src.slice_shape(k)
                    v = lb[:, sl]
                    dict.__setitem__(self, k, v if sh == () else v.reshape((lb.shape[0],) + sh))
                self._layout_from(src)
                self._buf = lb
                return
            if rbatch_buf is not None:
                rb = _np.asarray(rbatch_buf)
                for k in src:
                    sl, sh = src.slice_shape(k)
                    v = rb[sl]
                    dict.__setitem__(self, k, v if sh == () else v.reshape(sh + (rb.shape[-1],)))
                self._layout_from(src)
                self._buf = rb
                return
            b = _np.asarray(buf)
            if b.dtype != object:
                b = _np.array(b, dtype=float)
            self._layout_from(src)
            self._buf = b.reshape(-1)
            self._refresh()
            return
        items = src.items() if hasattr(src, 'keys') else src
        chunks = []
        n = 0
        for k, v in items:
            a = _np.asarray(v)
            if a.dtype != object:
                a = _np.array(a, dtype=float)
            if a.shape == ():
                self._slices[k] = (n, ())
                n += 1
            else:
                self._slices[k] = (slice(n, n + a.size), a.shape)
                n += a.size
            chunks.append(a.reshape(-1))
        if chunks:
            isobj = any(c.dtype == object for c in chunks)
            self._buf = _np.concatenate([c.astype(object) if isobj else c for c in chunks])
        else:
            self._buf = _np.zeros(0, float)
        self._refresh()

    def _layout_from(self, src):
        self._slices = dict(src._slices)
        for k in src:
            if k not in self:
                dict.__setitem__(self, k, None)

    def _refresh(self):
        for k, (sl, sh) in self._slices.items():
            dict.__setitem__(self, k, self._buf[sl] if sh == () else self._buf[sl].reshape(sh))

    def __setitem__(self, k, v):
        if k in self._slices:
            sl, sh = self._slices[k]
            if sh == ():
                self._buf[sl] = v
            else:
                self._buf[sl] = _np.asarray(v).reshape(-1)
            self._refresh()
            return
        a = _np.asarray(v)
        if a.dtype != object:
            a = _np.array(a, dtype=float)
        n = self._buf.size
        if a.shape == ():
            self._slices[k] = (n, ())
        else:
            self._slices[k] = (slice(n, n + a.size), a.shape)
        if a.dtype == object or self._buf.dtype == object:
            self._buf = _np.concatenate([self._buf.astype(object), a.reshape(-1).astype(object)])
        else:
            self._buf = _np.concatenate([self._buf, a.reshape(-1)])
        self._refresh()

    def slice_shape(self, k):
        return self._slices[k]

    # values are read through to the buffer on every access, so in-place updates of
    # ``buf`` (RAvgDict shares its buffer with a RAvgArray) are always visible
    def __getitem__(self, k):
        if k in self._slices and self._buf.ndim == 1:
            sl, sh = self._slices[k]
            return self._buf[sl] if sh == () else self._buf[sl].reshape(sh)
        return dict.__getitem__(self, k)

    def values(self):
        return [self[k] for k in self]

    def items(self):
        return [(k, self[k]) for k in self]

    def get(self, k, default=None):
        return self[k] if k in self else default

    def __repr__(self):
        return '{' + ', '.join('%r: %r' % (k, self[k]) for k in self) + '}'

    __str__ = __repr__

    def __truediv__(self, other):
        return BufferDict(self, buf=self.buf / other)

    def __mul__(self, other):
        return BufferDict(self, buf=self.buf * other)

    def slice(self, k):
        sl = self._slices[k][0]
        return sl if isinstance(sl, slice) else slice(sl, sl + 1)

    def all_keys(self):
        return list(self.keys())

    @property
    def buf(self):
        return self._buf

    @buf.setter
    def buf(self, b):
        b = _np.asarray(b) if not isinstance(b, _np.ndarray) else b
        self._buf = b.reshape(-1) if b.ndim != 1 else b
        self._refresh()

    @property
    def size(self):
        return self._buf.size if self._buf.ndim == 1 else sum(
            1 if sh == () else int(_np.prod(sh)) for _, sh in self._slices.values())

    @property
    def flat(self):
        return self._buf.flat

    def __reduce__(self):
        return (BufferDict, ([(k, self[k]) for k in self],))

    def _r2lbatch(self):
        """rbatch values (batch index last) -> (lbatch BufferDict, None); used by the reference's
        restratify auxiliary integrand (src/vegas/__init__.py:1419)"""
        out = BufferDict()
        for k in self:
            out[k] = _np.moveaxis(_np.asarray(self[k]), -1, 0)
        return out, None


def asbufferdict(g):
    return g if isinstance(g, BufferDict) else BufferDict(g)


class SVD(object):
    """Eigen-analysis of a symmetric matrix with gvar-style ``svdcut``.

    ``svdcut < 0`` drops modes whose eigenvalue is below ``|svdcut| * max``;
    ``svdcut > 0`` raises them to that floor.  ``decomp(n)`` returns rows
    ``w_i = vec_i * val_i**(n/2)`` so that ``sum_i w_i w_i^T = mat**n``.
    """

    def __init__(self, mat, svdcut=None, svdnum=None, compute_delta=False, rescale=False):
        mat = _np.asarray(mat, dtype=float)
        if rescale:
            self.D = _np.fabs(mat.diagonal()) ** (-0.5)
            work = (mat * self.D).T * self.D
        else:
            self.D = None
            work = mat
        vec, val, _ = _np.linalg.svd(work)
        vec = vec.T[::-1]
        val = val[::-1]                      # ascending
        self.kappa = val[0] / val[-1] if val[-1] != 0 else None
        self.nmod = 0
        self.dof = len(val)
        if svdcut:
            floor = abs(svdcut) * val[-1]
            if svdcut > 0:
                small = val < floor
                self.nmod = int(small.sum())
                val = _np.where(small, floor, val)
            else:
                keep = val >= floor
                first = int(_np.argmax(keep)) if keep.any() else len(val)
                val = val[first:]
                vec = vec[first:]
                self.dof = len(val)
        self.val = val
        self.vec = vec

    def decomp(self, n=1):
        w = _np.array(self.vec)
        if self.D is not None:
            w = w * (self.D if n < 0 else 1.0 / self.D)
        return (w.T * self.val ** (n / 2.0)).T


def dumps(g, protocol=None):
    return _pickle.dumps(g, protocol=protocol if protocol is not None else _pickle.HIGHEST_PROTOCOL)


def loads(b):
    return _pickle.loads(b)


def remove_gvars(g, gvlist):
    return g


def distribute_gvars(g, gvlist):
    return g


def gvar_factory():
    return gvar


def tabulate(g, **kargs):
    return '\n'.join('%s  %s' % (k, g[k]) for k in g) if hasattr(g, 'keys') else str(g)


class PDF(object):
    r""" Gaussian probability density of a collection of GVars (the part of ``gvar.PDF`` that
    ``PDFIntegrator`` uses, reference ``src/vegas/__init__.py:496-503, 599-606, 1163-1183``).

    The parameters are written ``p = mean + chiv . vec_sig`` with ``chiv`` a vector of independent
    unit normal variables along the principal axes of the correlation matrix (eigenvalues below
    ``svdcut`` times the largest are raised to that floor).  ``dp_dchiv`` is the Jacobian of the map.
    """

    def __init__(self, g, svdcut=1e-12):
        if isinstance(g, PDF):
            self.__dict__.update(g.__dict__)
            return
        if hasattr(g, 'keys'):
            self.g = asbufferdict(g)
            self.shape = None
            flat = _np.asarray(self.g.buf, dtype=object).reshape(-1)
        else:
            self.g = _np.asarray(g, dtype=object)
            self.shape = self.g.shape
            flat = self.g.reshape(-1)
        self.size = int(flat.size)
        self.svdcut = svdcut
        self.meanflat = _np.asarray(mean(flat), dtype=float).reshape(-1)
        cov = _np.asarray(evalcov(flat), dtype=float).reshape(self.size, self.size)
        if _np.any(cov.diagonal() <= 0):
            raise ValueError('PDF needs parameters with non-zero standard deviations')
        svd = SVD(cov, svdcut=abs(svdcut) if svdcut else None, rescale=True)
        self.vec_sig = svd.decomp(1)                 # rows w_i: cov = sum_i w_i w_i^T
        self.vec_isig = svd.decomp(-1)               # rows v_i: cov^-1 = sum_i v_i v_i^T
        self.dp_dchiv = float(_np.prod(svd.val ** 0.5) / _np.prod(svd.D))
        self.log_gnorm = float(-0.5 * self.size * _np.log(2 * _np.pi) - _np.log(self.dp_dchiv))

    # --- maps between the parameters and the unit-normal variables
    def pflat(self, chiv, mode=None):
        chiv = _np.asarray(chiv, dtype=float)
        if mode == 'rbatch':
            return self.meanflat[:, None] + self.vec_sig.T.dot(chiv)
        return self.meanflat + chiv.dot(self.vec_sig)            # None: chiv[i]; 'lbatch': chiv[n, i]

    def chiv(self, p, mode=None):
        pf = self._flatten(p, mode)
        if mode == 'rbatch':
            return self.vec_isig.dot(pf - self.meanflat[:, None])
        return (pf - self.meanflat).dot(self.vec_isig.T)

    def _flatten(self, p, mode=None):
        if hasattr(p, 'keys'):
            p = asbufferdict(p)
            b = _np.asarray(p.buf, dtype=float)
            return b
        p = _np.asarray(p, dtype=float)
        if mode == 'lbatch':
            return p.reshape(p.shape[0], -1)
        if mode == 'rbatch':
            return p.reshape(-1, p.shape[-1])
        return p.reshape(-1)

    def _unflatten(self, pflat, mode=None):
        pflat = _np.asarray(pflat)
        if self.shape is None:
            if mode == 'lbatch':
                return BufferDict(self.g, lbatch_buf=pflat)
            if mode == 'rbatch':
                return BufferDict(self.g, rbatch_buf=pflat)
            return BufferDict(self.g, buf=pflat)
        if mode == 'lbatch':
            return pflat.reshape((pflat.shape[0],) + self.shape)
        if mode == 'rbatch':
            return pflat.reshape(self.shape + (pflat.shape[-1],))
        return pflat.reshape(self.shape) if self.shape != () else pflat.reshape(-1)[0]

    def sample(self, nbatch=None, mode=None):
        """random parameter values drawn from the distribution, in the layout of ``g``"""
        if nbatch is None or mode is None:
            return self._unflatten(self.pflat(RNG.normal(size=self.size)))
        if mode == 'lbatch':
            return self._unflatten(self.pflat(RNG.normal(size=(nbatch, self.size)), mode='lbatch'), mode='lbatch')
        return self._unflatten(self.pflat(RNG.normal(size=(self.size, nbatch)), mode='rbatch'), mode='rbatch')

    def logpdf(self, p, mode=None):
        c = self.chiv(p, mode)
        return self.log_gnorm - 0.5 * _np.sum(c * c, axis=0 if mode == 'rbatch' else -1)

    def pdf(self, p, mode=None):
        return _np.exp(self.logpdf(p, mode))

    def __call__(self, p, mode=None):
        return self.pdf(p, mode)


class _Loc(object):
    """location with its lower and upper widths"""

    def __init__(self, loc, minus, plus):
        self.loc, self.minus, self.plus = loc, minus, plus

    def __str__(self):
        return '%s +/- %s/%s' % (self.loc, self.plus, self.minus)


class PDFStatistics(object):
    r""" Mean, standard deviation, skewness and excess kurtosis of a one-dimensional distribution from
    its moments ``[<x>, <x^2>, <x^3>, <x^4>]`` (GVars; the last two optional), and its median with the
    15.87 % / 84.13 % interval from a histogram ``(bins, count)`` (``count`` includes the underflow and
    overflow bins).  The real ``gvar.PDFStatistics`` also fits split normal distributions to the
    histogram; here ``splitnormal`` is the percentile-matched approximation (same object as ``median``).
    """

    def __init__(self, moments=None, histogram=None, prefix='   '):
        self.prefix = prefix
        self.mean = self.sdev = self.skew = self.ex_kurt = None
        self.median = self.splitnormal = None
        if moments is not None:
            mom = list(moments)
            self.mean = mom[0]
            var_ = mom[1] - mom[0] * mom[0]
            self.sdev = fabs(var_) ** 0.5
            if len(mom) > 2:
                self.skew = (mom[2] - 3. * self.mean * var_ - self.mean ** 3) / self.sdev ** 3
            if len(mom) > 3:
                m4 = mom[3] - 4. * mom[2] * self.mean + 6. * mom[1] * self.mean ** 2 - 3. * self.mean ** 4
                self.ex_kurt = m4 / (var_ * var_) - 3.
        if histogram is not None:
            bins, count = histogram
            self.bins = _np.asarray(bins, dtype=float)
            prob = count / _np.sum(count)
            self.prob = prob
            cum = _np.cumsum(prob[:-1])                  # probability below bins[i], i = 0..nbin

            def quantile(qv):
                cm = mean(cum)
                i = int(_np.searchsorted(cm, qv))
                i = min(max(i, 1), len(cm) - 1)
                x0, x1 = self.bins[i - 1], self.bins[i]
                c0, c1 = cum[i - 1], cum[i]
                return x0 + (x1 - x0) * ((qv - c0) / (c1 - c0))
            med = quantile(0.5)
            self.median = _Loc(med, med - quantile(0.158655253931457), quantile(0.841344746068543) - med)
            self.splitnormal = self.median

    def __str__(self):
        out = []
        if self.mean is not None:
            line = self.prefix + 'mean = %s   sdev = %s' % (self.mean, self.sdev)
            if self.skew is not None:
                line += '   skew = %s' % (self.skew,)
            if self.ex_kurt is not None:
                line += '   ex_kurt = %s' % (self.ex_kurt,)
            out.append(line)
        if self.median is not None:
            out.append(self.prefix + 'median: %s' % (self.median,))
        return '\n'.join(out)
