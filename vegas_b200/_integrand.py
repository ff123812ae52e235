"""Integrand adapters: every kind of user integrand is normalised to ``eval(x[n, D], jac) -> f[n, size]``
(the reference's ``VegasIntegrand`` and ``_BatchIntegrand_from_*`` classes, ``_vegas.pyx:2959-3383``),
plus the decorators (``_vegas.pyx:3386-3473``) and the two GPU-resident integrand kinds:

* ``DeviceIntegrand`` -- a functor compiled into ``libvegas_b200.so`` (fused path; see
  ``vegas_b200.integrands``);
* ``@devicebatchintegrand`` -- an lbatch callback that receives/returns CUDA ``torch`` tensors
  (DLPack-compatible), so samples never leave HBM (unfused path).
"""
import collections

import numpy as np

from ._gv import gv


# --------------------------------------------------------------------------- decorators / containers
class _BatchWrapper(object):
    """callable that marks a batch integrand: wraps ``fcn`` or -- subclassed with its own ``__call__`` -- is the
    integrand itself; unknown attributes are looked up on the wrapped function"""
    fcntype = None

    def __init__(self, fcn=None):
        self.fcn = self if fcn is None else fcn

    def __call__(self, *args, **kargs):
        return self.fcn(*args, **kargs)

    def __getattr__(self, attr):
        if attr == 'fcn' or self.fcn is self:
            raise AttributeError(attr)
        return getattr(self.fcn, attr)


class LBatchIntegrand(_BatchWrapper):
    r""" Wrapper for lbatch integrands (``x[i, d]``, batch index on the left). """
    fcntype = 'lbatch'


class RBatchIntegrand(_BatchWrapper):
    r""" Same as :class:`LBatchIntegrand` but with batch indices on the right (``x[d, i]``). """
    fcntype = 'rbatch'


def _tag(f, wrapper, **marks):
    """mark ``f`` itself when it accepts attributes (the meaning of ``f(x)`` is unchanged), else wrap it"""
    try:
        for k, v in marks.items():
            setattr(f, k, v)
        f.fcntype = wrapper.fcntype
        return f
    except Exception:
        w = wrapper(f)
        for k, v in marks.items():
            setattr(w, k, v)
        return w


def lbatchintegrand(f):
    r""" Decorator for batch integrand functions ``f(x[i, d]) -> f[i]`` (or arrays / dicts with a
    leading batch index).  The meaning of ``f(x)`` is unchanged. """
    return _tag(f, LBatchIntegrand)


def rbatchintegrand(f):
    r""" Same as :func:`lbatchintegrand` but with batch indices on the right. """
    return _tag(f, RBatchIntegrand)


def devicebatchintegrand(f):
    r""" Decorator for lbatch integrands evaluated on the GPU: ``f(x)`` receives a float64 CUDA
    ``torch.Tensor`` ``x[i, d]`` (DLPack-exportable: ``cupy.from_dlpack(x)``, ``jax.dlpack`` ...)
    and returns a CUDA tensor (or any object exporting ``__dlpack__``) ``f[i]`` or ``f[i, ...]``.
    Samples and integrand values then stay in HBM (the unfused device-callback path). """
    return _tag(f, LBatchIntegrand, on_device=True)


# legacy names (reference _vegas.pyx:3465-3473)
batchintegrand = lbatchintegrand
BatchIntegrand = LBatchIntegrand
vecintegrand = batchintegrand
MPIintegrand = batchintegrand


class VecIntegrand(LBatchIntegrand):
    pass


class DeviceIntegrand(object):
    r""" An integrand whose device functor is compiled into the library.

    Subclasses (``vegas_b200.integrands``) provide ``fid``, ``params(dim)`` (a ctypes struct for
    ``vb200_set_integrand``) and a numpy twin ``__call__(x[n, D]) -> f[n]`` / ``f[n, nf]`` with the same
    constants, so the same object also works as an ordinary lbatch integrand (e.g. on the CPU
    reference)."""
    fcntype = 'lbatch'
    fid = None
    nf = 1
    shape = ()            # shape of the integrand's value; () = scalar, None = dict (then ``keys``)

    def params(self, dim):
        raise NotImplementedError

    def format(self, buf):
        """repack the flat [nf] result (objects) into the integrand's output structure"""
        return buf

    def device_twin(self, dim, device=None):
        """This integrand as a ``@devicebatchintegrand``: ``f(x)`` runs the library functor on the HBM
        buffer ``x[n, dim]`` (``vb200_eval_integrand``) and returns ``f[n]`` / ``f[n, nf]`` in HBM -- the
        device batch callback route of ``Integrator.__call__`` (sample -> callback -> reduce) with the
        same arithmetic as the fused kernel."""
        import torch
        from . import _lib
        ctx = _lib.Context(device)
        ctx.set_map(np.tile([0., 1.], (dim, 1)), np.ones(dim, np.int64))      # the functor only needs the dimension
        params, keep = self.params(dim)
        nf = ctx.set_integrand(self.fid, params, keep=(params, keep))

        def f(x):
            out = torch.empty((x.shape[0], nf), dtype=torch.float64, device=x.device)
            ctx.eval_integrand(x.contiguous(), out)
            return out if nf > 1 else out[:, 0]
        f._ctx = ctx
        return devicebatchintegrand(f)


# --------------------------------------------------------------------------- standard form
# Every host integrand becomes  eval(x[n, D], jac[n, D] | None) -> f[n, size]  by composing two pieces, both
# chosen once from the probe call (the reference has one class per combination, _vegas.pyx:3175-3383):
#   presenter  (_Presenter) rows -> the argument the user function takes: flat rows, an index array of xsample's shape, or a
#              dictionary; one point at a time, or a batch with its index on the left / on the right.  jac is
#              presented exactly like x.
#   packer     the user function's value (number, array or dictionary; per point or per batch) -> rows of `size`.
class _Presenter(object):
    """rows ``a[n, D]`` (``a[D]`` for one point) -> the argument form of the user function"""

    def __init__(self, xsample, kind):
        self.xsample, self.kind = xsample, kind
        self.shape = None if xsample.shape is None else tuple(xsample.shape)
        if self.shape is None:
            self.how = 'dict'
        elif len(self.shape) == 1:
            self.how = 'flat'
        else:
            self.how = 'index'

    def __call__(self, a):
        right = self.kind == 'rbatch'
        if self.how == 'flat':
            return a.T if right else a
        if self.how == 'dict':
            if right:
                return gv.BufferDict(self.xsample, rbatch_buf=a.T)
            if self.kind == 'lbatch':
                return gv.BufferDict(self.xsample, lbatch_buf=a)
            return gv.BufferDict(self.xsample, buf=a)
        if right:
            return a.T.reshape(self.shape + (-1,))
        return a.reshape(((-1,) if self.kind == 'lbatch' else ()) + self.shape)


def _dict_layout(bdict):
    """[(key, slice or index into the flat row, shape)] of a BufferDict"""
    return [(k,) + tuple(bdict.slice_shape(k)) for k in bdict]


class _HostEval(object):
    """``eval`` of a host (numpy) integrand.  ``kind``: 'scalar' | 'lbatch' | 'rbatch'; ``shape``: the value's own
    shape, or None for a dictionary (then ``bdict`` is a sample of it)."""

    def __init__(self, fcn, xsample, kind, shape, size, bdict):
        self.fcn, self.kind, self.shape, self.size = fcn, kind, shape, int(size)
        self.present = _Presenter(xsample, kind)
        self.layout = None if bdict is None else _dict_layout(bdict)
        # how jac reaches a one-point function (pyx:3200-3213): by keyword for flat x, positionally for a
        # dictionary, not at all for an index array
        self.point_jac = 'second' if xsample.shape is None else ('keyword' if len(xsample.shape) == 1 else 'dropped')

    def _point(self, xrow, jrow):
        xa = self.present(xrow)
        if jrow is None or self.point_jac == 'dropped':
            return self.fcn(xa)
        if self.point_jac == 'second':
            return self.fcn(xa, self.present(jrow))
        return self.fcn(xa, jac=self.present(jrow))

    def __call__(self, x, jac=None):
        x = np.asarray(x)
        n = x.shape[0]
        if self.kind == 'scalar':
            out = np.empty((n, self.size), float)
            for i in range(n):
                fx = self._point(x[i], None if jac is None else np.asarray(jac[i]))
                if self.layout is not None:
                    out[i] = (fx if isinstance(fx, gv.BufferDict) else gv.BufferDict(fx)).buf[:self.size]
                elif self.shape == ():
                    out[i, 0] = fx
                else:
                    out[i] = np.asarray(fx).reshape(-1)
            return out
        fx = self.fcn(self.present(x)) if jac is None else self.fcn(self.present(x), jac=self.present(np.asarray(jac)))
        right = self.kind == 'rbatch'
        if self.layout is None:
            fx = np.asarray(fx)
            return np.ascontiguousarray(fx.reshape((-1, n)).T) if right else fx.reshape((n, -1))
        out = np.empty((n, self.size), float)
        for k, where, shape in self.layout:
            v = fx[k]
            if shape != ():
                v = np.reshape(v, (-1, n)).T if right else np.asarray(v).reshape((n, -1))
            out[:, where] = v
        return out


class _DeviceEval(object):
    """``eval`` of a ``@devicebatchintegrand``: torch CUDA tensors in, [n, size] tensor out"""

    def __init__(self, fcn):
        self.fcn = fcn

    def __call__(self, x, jac=None):
        import torch
        fx = self.fcn(x) if jac is None else self.fcn(x, jac=jac)
        if not isinstance(fx, torch.Tensor):
            fx = torch.from_dlpack(fx)
        if fx.dtype != torch.float64:
            fx = fx.to(torch.float64)
        return fx.reshape(x.shape[0], -1).contiguous()


def _strip_batch(value, kind):
    """one point's value out of a one-point batch"""
    v = np.asarray(value)
    return v[..., 0] if kind == 'rbatch' else v[0]


class VegasIntegrand(object):
    r""" Integrand object --- standard interface for integrands (``_vegas.pyx:2959-3169``).

    Analyzes ``fcn`` with one probe call on ``xsample`` to learn the shape of its output, then
    exposes ``eval(x[i, d], jac=None) -> f[i, c]``.

    Attributes: ``eval``, ``shape`` (``None`` for dictionaries), ``size``, ``fcntype``, ``bdict``,
    ``on_device`` (True for ``@devicebatchintegrand``), ``mpi_nproc``, ``rank``.
    """

    def __init__(self, fcn, map, uses_jac, xsample, mpi):
        if isinstance(fcn, type) and issubclass(fcn, _BatchWrapper):
            raise ValueError('integrand given is a class, not an object -- need to initialize?')
        self.mpi_nproc, self.rank = 1, 0
        self.on_device = bool(getattr(fcn, 'on_device', False))
        self.fcntype = getattr(fcn, 'fcntype', 'scalar')
        self.bdict = None
        xsample = gv.mean(xsample)
        x1 = np.array(xsample.buf if xsample.shape is None else xsample, dtype=float).reshape(1, -1)   # one row: the probe
        if self.on_device:
            import torch
            xd = torch.from_numpy(x1).cuda()
            fx = fcn(xd, jac=torch.ones_like(xd)) if uses_jac else fcn(xd)
            if not isinstance(fx, torch.Tensor):
                fx = torch.from_dlpack(fx)
            self.shape = tuple(fx.shape[1:])
            self.size = int(np.prod(self.shape, dtype=np.int64))
            self.eval = _DeviceEval(fcn)
            return
        kind = self.fcntype if self.fcntype in ('scalar', 'rbatch') else 'lbatch'
        present = _Presenter(xsample, kind)
        if kind == 'scalar':
            # (the probe always passes jac by keyword, as pyx:3013 does; eval follows pyx:3200-3213)
            fx = fcn(present(x1[0]), jac=present(np.ones_like(x1[0]))) if uses_jac else fcn(present(x1[0]))
            if hasattr(fx, 'keys'):
                self.bdict = fx if isinstance(fx, gv.BufferDict) else gv.BufferDict(fx)
                self.shape, self.size = None, self.bdict.size
            else:
                fx = np.asarray(fx)
                self.shape, self.size = fx.shape, fx.size
        else:
            fx = fcn(present(x1), jac=present(np.ones_like(x1))) if uses_jac else fcn(present(x1))
            if hasattr(fx, 'keys'):
                self.bdict = gv.BufferDict()
                for k in fx:
                    self.bdict[k] = _strip_batch(fx[k], kind)
                self.shape, self.size = None, self.bdict.size
            else:
                self.shape = tuple(np.shape(fx)[:-1] if kind == 'rbatch' else np.shape(fx)[1:])
                self.size = int(np.prod(self.shape, dtype=np.int64))
        self.eval = _HostEval(fcn, xsample, kind, self.shape, self.size, self.bdict)

    def __call__(self, x, jac=None):
        r""" Non-batch version of fcn """
        row = gv.asbufferdict(x).buf if hasattr(x, 'keys') else np.asarray(x)
        return self.format_result(np.asarray(self.eval(row.reshape(1, -1), jac=jac)))

    def _structured(self, flat):
        """a flat vector of ``size`` entries in the integrand's own structure"""
        if self.shape is None:
            return gv.BufferDict(self.bdict, buf=flat.reshape(-1))
        return flat.reshape(self.shape)

    def format_result(self, mean, var=None):
        r""" Reformat output from integrator to correspond to original output format: ``mean`` alone, or GVars
        from ``mean`` and ``var`` (a vector of variances or a covariance matrix). """
        if var is None:
            return mean.flat[0] if self.shape == () else self._structured(mean)
        if var.shape == mean.shape:
            spread = np.asarray(var) ** 0.5
            if self.shape == ():
                return gv.gvar(mean[0], spread[0])
        else:
            spread = var
            if self.shape == ():
                return gv.gvar(mean[0], var[0, 0] ** 0.5)
        return self._structured(gv.gvar(mean, spread))

    def format_evalx(self, evalx):
        r""" Reformat output ``evalx[i, c]`` of ``eval(x)`` into the integrand's own structure. """
        if self.shape is None:
            return gv.BufferDict(self.bdict, lbatch_buf=evalx)
        return evalx.reshape(evalx.shape[:1] + self.shape)

    def training(self, x, jac):
        r""" Calculate first element of integrand at point ``x``. """
        fx = self.eval(x, jac=jac)
        return fx if fx.ndim == 1 else fx.reshape((x.shape[0], -1))[:, 0]
