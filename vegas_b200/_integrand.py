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
class LBatchIntegrand(object):
    r""" Wrapper for lbatch integrands (``x[i, d]``, batch index on the left). """

    def __init__(self, fcn=None):
        self.fcn = self if fcn is None else fcn

    fcntype = 'lbatch'

    def __call__(self, *args, **kargs):
        return self.fcn(*args, **kargs)

    def __getattr__(self, attr):
        if attr == 'fcn' or self.fcn is self:
            raise AttributeError(attr)
        return getattr(self.fcn, attr)


class RBatchIntegrand(object):
    r""" Same as :class:`LBatchIntegrand` but with batch indices on the right (``x[d, i]``). """

    def __init__(self, fcn=None):
        self.fcn = self if fcn is None else fcn

    fcntype = 'rbatch'

    def __call__(self, *args, **kargs):
        return self.fcn(*args, **kargs)

    def __getattr__(self, attr):
        if attr == 'fcn' or self.fcn is self:
            raise AttributeError(attr)
        return getattr(self.fcn, attr)


def lbatchintegrand(f):
    r""" Decorator for batch integrand functions ``f(x[i, d]) -> f[i]`` (or arrays / dicts with a
    leading batch index).  The meaning of ``f(x)`` is unchanged. """
    try:
        f.fcntype = 'lbatch'
        return f
    except Exception:
        return LBatchIntegrand(f)


def rbatchintegrand(f):
    r""" Same as :func:`lbatchintegrand` but with batch indices on the right. """
    try:
        f.fcntype = 'rbatch'
        return f
    except Exception:
        return RBatchIntegrand(f)


def devicebatchintegrand(f):
    r""" Decorator for lbatch integrands evaluated on the GPU: ``f(x)`` receives a float64 CUDA
    ``torch.Tensor`` ``x[i, d]`` (DLPack-exportable: ``cupy.from_dlpack(x)``, ``jax.dlpack`` ...)
    and returns a CUDA tensor (or any object exporting ``__dlpack__``) ``f[i]`` or ``f[i, ...]``.
    Samples and integrand values then stay in HBM (the unfused device-callback path). """
    try:
        f.fcntype = 'lbatch'
        f.on_device = True
        return f
    except Exception:
        w = LBatchIntegrand(f)
        w.on_device = True
        return w


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
class _Base(object):
    """manages xsample: how flat x rows are presented to the user function"""

    def __init__(self, fcn, xsample):
        self.fcn = fcn
        self.xsample = xsample
        if xsample.shape is None:
            self.dict_arg, self.std_arg = True, False
        else:
            self.dict_arg, self.std_arg = False, len(xsample.shape) == 1

    def _one(self, x, jac=None):
        " fcn(x) for one point when the argument is a dict or a multi-index array "
        x = np.asarray(x)
        if self.dict_arg:
            xd = gv.BufferDict(self.xsample, buf=x)
            if jac is not None:
                return self.fcn(xd, gv.BufferDict(self.xsample, buf=jac))
            return self.fcn(xd)
        return self.fcn(x.reshape(self.xsample.shape))

    def _batch(self, x, jac=None):
        " fcn(x) for a batch when the argument is a dict or a multi-index array "
        x = np.asarray(x)
        if self.dict_arg:
            if self.rbatch:
                xd = gv.BufferDict(self.xsample, rbatch_buf=x.T)
                if jac is not None:
                    jac = gv.BufferDict(self.xsample, rbatch_buf=jac.T)
            else:
                xd = gv.BufferDict(self.xsample, lbatch_buf=x)
                if jac is not None:
                    jac = gv.BufferDict(self.xsample, lbatch_buf=jac)
            return self.fcn(xd) if jac is None else self.fcn(xd, jac=jac)
        if self.rbatch:
            sh = self.xsample.shape + (-1,)
            return self.fcn(x.T.reshape(sh)) if jac is None else self.fcn(x.T.reshape(sh), jac=jac.T.reshape(sh))
        sh = (-1,) + self.xsample.shape
        return self.fcn(x.reshape(sh)) if jac is None else self.fcn(x.reshape(sh), jac=jac.reshape(sh))


class _FromNonBatch(_Base):
    """ batch integrand from a scalar (one point at a time) integrand """

    def __init__(self, fcn, size, shape, xsample):
        self.size, self.shape = size, shape
        _Base.__init__(self, fcn, xsample)

    def __call__(self, x, jac=None):
        x = np.asarray(x)
        f = np.empty((x.shape[0], self.size), float)
        for i in range(x.shape[0]):
            ji = None if jac is None else jac[i]
            if self.std_arg:
                fx = self.fcn(x[i]) if ji is None else self.fcn(x[i], jac=ji)
            else:
                fx = self._one(x[i], ji)
            if self.shape == ():
                f[i, 0] = fx
            else:
                f[i] = np.asarray(fx).reshape(-1)
        return f


class _FromNonBatchDict(_Base):
    """ batch integrand from a scalar integrand that returns a dictionary """

    def __init__(self, fcn, size, xsample):
        self.size = size
        _Base.__init__(self, fcn, xsample)

    def __call__(self, x, jac=None):
        x = np.asarray(x)
        f = np.empty((x.shape[0], self.size), float)
        for i in range(x.shape[0]):
            ji = None if jac is None else jac[i]
            if self.std_arg:
                fx = self.fcn(x[i]) if ji is None else self.fcn(x[i], jac=ji)
            else:
                fx = self._one(x[i], ji)
            if not isinstance(fx, gv.BufferDict):
                fx = gv.BufferDict(fx)
            f[i] = fx.buf[:self.size]
        return f


class _FromBatch(_Base):
    """ standard form of an lbatch / rbatch integrand returning arrays """

    def __init__(self, fcn, rbatch, xsample):
        self.rbatch = rbatch
        _Base.__init__(self, fcn, xsample)

    def __call__(self, x, jac=None):
        if self.std_arg:
            if self.rbatch:
                fx = self.fcn(x.T) if jac is None else self.fcn(x.T, jac=jac.T)
            else:
                fx = self.fcn(x) if jac is None else self.fcn(x, jac=jac)
        else:
            fx = self._batch(x, jac)
        if not isinstance(fx, np.ndarray):
            fx = np.asarray(fx)
        if self.rbatch:
            return np.ascontiguousarray(fx.reshape((-1, x.shape[0])).T)
        return fx.reshape((x.shape[0], -1))


class _FromBatchDict(_Base):
    """ standard form of an lbatch / rbatch integrand returning a dictionary """

    def __init__(self, fcn, bdict, rbatch, xsample):
        self.size = bdict.size
        self.rbatch = rbatch
        self.slice = collections.OrderedDict()
        self.shape = collections.OrderedDict()
        for k in bdict:
            self.slice[k], self.shape[k] = bdict.slice_shape(k)
        _Base.__init__(self, fcn, xsample)

    def __call__(self, x, jac=None):
        buf = np.empty((x.shape[0], self.size), float)
        if self.std_arg:
            if self.rbatch:
                fx = self.fcn(x.T) if jac is None else self.fcn(x.T, jac=jac.T)
            else:
                fx = self.fcn(x) if jac is None else self.fcn(x, jac=jac)
        else:
            fx = self._batch(x, jac)
        for k in self.slice:
            if self.shape[k] == ():
                buf[:, self.slice[k]] = fx[k]
            elif self.rbatch:
                buf[:, self.slice[k]] = np.reshape(fx[k], (-1, x.shape[0])).T
            else:
                buf[:, self.slice[k]] = np.asarray(fx[k]).reshape((x.shape[0], -1))
        return buf


class _FromDeviceBatch(object):
    """ standard form of a ``@devicebatchintegrand``: torch CUDA tensors in, [n, size] tensor out """

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


class VegasIntegrand(object):
    r""" Integrand object --- standard interface for integrands (``_vegas.pyx:2959-3169``).

    Analyzes ``fcn`` with one probe call on ``xsample`` to learn the shape of its output, then
    exposes ``eval(x[i, d], jac=None) -> f[i, c]``.

    Attributes: ``eval``, ``shape`` (``None`` for dictionaries), ``size``, ``fcntype``, ``bdict``,
    ``on_device`` (True for ``@devicebatchintegrand``), ``mpi_nproc``, ``rank``.
    """

    def __init__(self, fcn, map, uses_jac, xsample, mpi):
        if isinstance(fcn, type) and issubclass(fcn, (LBatchIntegrand, RBatchIntegrand)):
            raise ValueError('integrand given is a class, not an object -- need to initialize?')
        self.mpi_nproc, self.rank = 1, 0
        self.on_device = bool(getattr(fcn, 'on_device', False))
        self.bdict = None
        xsample = gv.mean(xsample)
        x0 = xsample
        if uses_jac:
            if xsample.shape is None:
                jac0 = gv.BufferDict(xsample, buf=xsample.size * [1])
            else:
                jac0 = np.ones(xsample.shape, dtype=float)
        else:
            jac0 = None
        self.fcntype = getattr(fcn, 'fcntype', 'scalar')
        if self.on_device:
            import torch
            xs = np.asarray(xsample.buf if xsample.shape is None else xsample, dtype=float).reshape(1, -1)
            xd = torch.from_numpy(xs).cuda()
            jd = torch.ones_like(xd) if uses_jac else None
            fx = fcn(xd, jac=jd) if uses_jac else fcn(xd)
            if not isinstance(fx, torch.Tensor):
                fx = torch.from_dlpack(fx)
            self.shape = tuple(fx.shape[1:])
            self.size = int(np.prod(self.shape, dtype=np.int64))
            self.eval = _FromDeviceBatch(fcn)
        elif self.fcntype == 'scalar':
            fx = fcn(x0, jac=jac0) if uses_jac else fcn(x0)
            if hasattr(fx, 'keys'):
                if not isinstance(fx, gv.BufferDict):
                    fx = gv.BufferDict(fx)
                self.size, self.shape, self.bdict = fx.size, None, fx
                self.eval = _FromNonBatchDict(fcn, self.size, xsample)
            else:
                fx = np.asarray(fx)
                self.shape, self.size = fx.shape, fx.size
                self.eval = _FromNonBatch(fcn, self.size, self.shape, xsample)
        elif self.fcntype == 'rbatch':
            if x0.shape is None:
                x0 = gv.BufferDict(x0, rbatch_buf=x0.buf.reshape(x0.buf.shape + (1,)))
                if uses_jac:
                    jac0 = gv.BufferDict(jac0, rbatch_buf=jac0.buf.reshape(jac0.buf.shape + (1,)))
            else:
                x0 = x0.reshape(x0.shape + (1,))
                if uses_jac:
                    jac0 = jac0.reshape(jac0.shape + (1,))
            fx = fcn(x0, jac=jac0) if uses_jac else fcn(x0)
            if hasattr(fx, 'keys'):
                fxs = gv.BufferDict()
                for k in fx:
                    fxs[k] = np.asarray(fx[k])[..., 0]
                self.shape, self.bdict, self.size = None, fxs, fxs.size
                self.eval = _FromBatchDict(fcn, self.bdict, True, xsample)
            else:
                self.shape = np.shape(fx)[:-1]
                self.size = int(np.prod(self.shape, dtype=np.int64))
                self.eval = _FromBatch(fcn, True, xsample)
        else:
            if x0.shape is None:
                x0 = gv.BufferDict(x0, lbatch_buf=x0.buf.reshape((1,) + x0.buf.shape))
                if uses_jac:
                    jac0 = gv.BufferDict(jac0, lbatch_buf=jac0.buf.reshape((1,) + jac0.buf.shape))
            else:
                x0 = x0.reshape((1,) + x0.shape)
                if uses_jac:
                    jac0 = jac0.reshape((1,) + jac0.shape)
            fx = fcn(x0) if jac0 is None else fcn(x0, jac=jac0)
            if hasattr(fx, 'keys'):
                fxs = gv.BufferDict()
                for k in fx:
                    fxs[k] = np.asarray(fx[k])[0]
                self.shape, self.bdict, self.size = None, fxs, fxs.size
                self.eval = _FromBatchDict(fcn, self.bdict, False, xsample)
            else:
                fx = np.asarray(fx)
                self.shape = fx.shape[1:]
                self.size = int(np.prod(self.shape, dtype=np.int64))
                self.eval = _FromBatch(fcn, False, xsample)

    def __call__(self, x, jac=None):
        r""" Non-batch version of fcn """
        if hasattr(x, 'keys'):
            x = gv.asbufferdict(x).buf.reshape(1, -1)
        else:
            x = np.asarray(x).reshape(1, -1)
        return self.format_result(np.asarray(self.eval(x, jac=jac)))

    def format_result(self, mean, var=None):
        r""" Reformat output from integrator to correspond to original output format """
        if var is None:
            if self.shape is None:
                return gv.BufferDict(self.bdict, buf=mean.reshape(-1))
            if self.shape == ():
                return mean.flat[0]
            return mean.reshape(self.shape)
        if var.shape == mean.shape:
            var = np.asarray(var) ** 0.5
        if self.shape is None:
            return gv.BufferDict(self.bdict, buf=gv.gvar(mean, var).reshape(-1))
        if self.shape == ():
            return gv.gvar(mean[0], var[0, 0] ** 0.5 if var.shape != mean.shape else var[0])
        return gv.gvar(mean, var).reshape(self.shape)

    def format_evalx(self, evalx):
        r""" Reformat output ``evalx[i, c]`` of ``eval(x)`` into the integrand's own structure. """
        if self.shape is None:
            return gv.BufferDict(self.bdict, lbatch_buf=evalx)
        return evalx.reshape(evalx.shape[:1] + self.shape)

    def training(self, x, jac):
        r""" Calculate first element of integrand at point ``x``. """
        fx = self.eval(x, jac=jac)
        if fx.ndim == 1:
            return fx
        return fx.reshape((x.shape[0], -1))[:, 0]
