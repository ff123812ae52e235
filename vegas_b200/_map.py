"""``AdaptiveMap`` -- same interface as the reference's (``_vegas.pyx:39-832``), B200 engine inside.

State (``grid``, ``inc``, ``ninc``, training sums) is small and lives on the host as numpy arrays;
the array methods ``map`` / ``invmap`` / ``jac1d`` / ``add_training_data`` run as CUDA kernels
through the C ABI (``vb200_map`` ...), ``adapt`` is the library's host routine ``vb200_map_adapt``.
"""
import numpy as np

from . import _lib
from ._gv import gv

TINY = 10.0 ** -257                                                             # _vegas.pyx:34 (10 ** (min_10_exp + 50))
HUGE = 10.0 ** 258                                                              # _vegas.pyx:35
EPSILON = np.finfo(float).eps * 1e4                                             # _vegas.pyx:36


def _torch():
    import torch
    return torch


class AdaptiveMap(object):
    r"""Adaptive map ``y -> x(y)`` for multidimensional ``y`` and ``x`` (reference docstring at
    ``_vegas.pyx:39-110``).  ``grid[d][i]`` is the ``i``-th node in direction ``d``; ``ninc`` optionally
    regrids to a different number of increments with the same Jacobian."""

    # ``grid`` and ``inc`` are host arrays.  After a device-side adapt (``Integrator`` fast path:
    # ``vb200_map_adapt_device``) the current grid lives in the integrator's context only; it is copied back
    # the first time somebody looks (``_device_owner`` = that context until then).
    _device_owner = None

    def _pull(self):
        ctx, self._device_owner = self._device_owner, None
        if ctx is not None:
            g = ctx.get_map(self._grid.shape)
            inc = np.zeros_like(self._inc)
            for d in range(g.shape[0]):
                n = int(self.ninc[d])
                inc[d, :n] = g[d, 1:n + 1] - g[d, :n]
            self._grid, self._inc = g, inc

    @property
    def grid(self):
        if self._device_owner is not None:
            self._pull()
        return self._grid

    @grid.setter
    def grid(self, value):
        self._device_owner = None
        self._grid = value

    @property
    def inc(self):
        if self._device_owner is not None:
            self._pull()
        return self._inc

    @inc.setter
    def inc(self, value):
        self._inc = value

    def _adapted_on_device(self, ctx):
        """the integrator's context holds a newer grid than the host arrays (same shape, same ninc)"""
        self._device_owner = ctx
        self.clear()
        self._changed()

    def __init__(self, grid, ninc=None):
        self._ctx = None
        self._ctx_version = -1
        self._version = 0
        if isinstance(grid, AdaptiveMap):
            self.ninc = np.array(grid.ninc)
            self.inc = np.array(grid.inc)
            self.grid = np.array(grid.grid)
        else:
            dim = len(grid)
            len_g = np.array([len(x) for x in grid], dtype=np.intp)
            if min(len_g) < 2:
                raise ValueError('grid[d] must have at least 2 elements, not {}'.format(min(len_g)))
            self.ninc = len_g - 1
            self.inc = np.empty((dim, max(len_g) - 1), float)
            self.grid = np.empty((dim, self.inc.shape[1] + 1), float)
            for d in range(dim):
                nodes = sorted(float(v) for v in grid[d])
                self.grid[d, :len(nodes)] = nodes
                self.grid[d, len(nodes):] = nodes[-1]
                self.inc[d, :len(nodes) - 1] = self.grid[d, 1:len(nodes)] - self.grid[d, :len(nodes) - 1]
                self.inc[d, len(nodes) - 1:] = 0.0
        self.clear()
        if ninc is not None and not np.all(ninc == self.ninc):
            if np.all(np.asarray(self.ninc) == 1):
                self.make_uniform(ninc=ninc)
            else:
                self.adapt(ninc=ninc)

    # ------------------------------------------------------------------ simple accessors
    @property
    def dim(self):
        " Number of dimensions."
        return self._grid.shape[0]         # (the shape only: no copy back of a device-adapted grid)

    def region(self, d=-1):
        r""" x-space region: ``(xl, xu)`` for direction ``d``, or the list for all directions. """
        if d < 0:
            return [self.region(d) for d in range(self.dim)]
        return (self.grid[d, 0], self.grid[d, self.ninc[d]])

    def extract_grid(self):
        " Return a list of lists specifying the map's grid. "
        return [list(self.grid[d, :self.ninc[d] + 1]) for d in range(self.dim)]

    def __reduce__(self):
        return (AdaptiveMap, (self.extract_grid(),))

    def settings(self, ngrid=5):
        r""" Create string with information about grid nodes (at most ``ngrid`` per direction). """
        ans = []
        if ngrid > 0:
            for d in range(self.dim):
                grid_d = np.array(self.grid[d, :self.ninc[d] + 1])
                nskip = int(self.ninc[d] // ngrid)
                if nskip < 1:
                    nskip = 1
                start = nskip // 2
                ans += ["    grid[%2d] = %s" % (
                    d, np.array2string(grid_d[start::nskip], precision=3, prefix='    grid[xx] = '))]
        return '\n'.join(ans) + '\n'

    def random(self, n=None):
        " Create ``n`` random points in |y| space. "
        y = gv.RNG.random(self.dim) if n is None else gv.RNG.random((n, self.dim))
        return self(y)

    def clear(self):
        " Clear information accumulated by :meth:`AdaptiveMap.add_training_data`. "
        self.sum_f = None
        self.n_f = None

    def _changed(self):
        self._version += 1

    def _ninc_arg(self, ninc, err):
        if ninc is None:
            return np.array(self.ninc, dtype=np.intp)
        if np.shape(ninc) == ():
            return np.full(self.dim, int(ninc), dtype=np.intp)
        if np.shape(ninc) == (self.dim,):
            return np.array(ninc, dtype=np.intp)
        raise ValueError(err.format(np.shape(ninc) if 'shape' in err else str(ninc)))

    def make_uniform(self, ninc=None):
        r""" Replace the grid with a uniform grid (``_vegas.pyx:205-241``). """
        ninc = self._ninc_arg(ninc, 'ninc has wrong shape -- {}')
        if min(ninc) < 1:
            raise ValueError("no of increments < 1 in AdaptiveMap -- %s" % str(ninc))
        new_inc = np.zeros((self.dim, max(ninc)), float)
        new_grid = np.empty((self.dim, new_inc.shape[1] + 1), float)
        for d in range(self.dim):
            tmp = np.linspace(self.grid[d, 0], self.grid[d, self.ninc[d]], ninc[d] + 1)
            new_grid[d, :ninc[d] + 1] = tmp
            new_grid[d, ninc[d] + 1:] = tmp[-1]
            new_inc[d, :ninc[d]] = new_grid[d, 1:ninc[d] + 1] - new_grid[d, :ninc[d]]
        self.ninc, self.grid, self.inc = ninc, new_grid, new_inc
        self.clear()
        self._changed()

    # ------------------------------------------------------------------ device plumbing
    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.Context()
        if self._ctx_version != self._version:
            self._ctx.set_map(self.grid, self.ninc)
            self._ctx_version = self._version
        return self._ctx

    def _dev(self, a, ny, cols=True):
        """(device tensor view [ny, dim] or [ny], was_host) for a numpy array or torch tensor"""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            if not a.is_cuda or a.dtype != torch.float64 or not a.is_contiguous():
                raise ValueError('device arrays must be contiguous float64 CUDA tensors')
            return a[:ny], False
        a = np.ascontiguousarray(np.asarray(a, dtype=float)[:ny])
        return torch.from_numpy(a).to(self._context().device), True

    def __call__(self, y):
        r""" Return ``x`` values corresponding to ``y`` (single point or array ``y[..., d]``). """
        if y is None:
            y = gv.RNG.random(size=self.dim)
        else:
            y = np.array(y, float)
        y_shape = y.shape
        y = y.reshape(-1, y.shape[-1])
        x = np.empty_like(y)
        jac = np.empty(y.shape[0], float)
        self.map(y, x, jac)
        return x.reshape(y_shape)

    def jac1d(self, y):
        r""" One-dimensional Jacobians ``dx[d]/dy[d]`` at ``y[..., d]`` (``_vegas.pyx:265-295``). """
        torch = _torch()
        if isinstance(y, torch.Tensor):
            out = torch.empty_like(y)
            self._context().jac1d(y.reshape(-1, self.dim), out)
            return out
        y = np.asarray(y, float)
        y_shape = y.shape
        yd, _ = self._dev(y.reshape(-1, y_shape[-1]), None)
        out = torch.empty_like(yd)
        self._context().jac1d(yd, out)
        return out.cpu().numpy().reshape(y_shape)

    def jac(self, y):
        r""" Multidimensional Jacobian ``dx/dy`` at ``y[..., d]``. """
        return np.prod(self.jac1d(y), axis=-1)

    def map(self, y, x, jac, ny=-1):
        r""" Map ``y`` to ``x``: fills ``x[i, d]`` and ``jac[i]`` for ``i < ny`` (``_vegas.pyx:310-360``).

        Arguments are numpy arrays (copied through the GPU) or contiguous float64 CUDA tensors
        (used in place)."""
        if ny < 0:
            ny = y.shape[0]
        elif ny > y.shape[0]:
            raise ValueError('ny > y.shape[0]: %d > %d' % (ny, y.shape[0]))
        torch = _torch()
        ctx = self._context()
        yd, host = self._dev(y, ny)
        if host:
            xd = torch.empty_like(yd)
            jd = torch.empty(ny, dtype=torch.float64, device=yd.device)
            ctx.map(yd, xd, jd)
            np.asarray(x)[:ny] = xd.cpu().numpy()
            np.asarray(jac)[:ny] = jd.cpu().numpy()
        else:
            ctx.map(yd, x, jac)

    def invmap(self, x, y, jac, nx=-1):
        r""" Map ``x`` to ``y`` (inverse map): fills ``y[i, d]`` and ``jac[i]`` (``_vegas.pyx:362-416``). """
        if nx < 0:
            nx = x.shape[0]
        elif nx > x.shape[0]:
            raise ValueError('nx > x.shape[0]: %d > %d' % (nx, x.shape[0]))
        torch = _torch()
        ctx = self._context()
        xd, host = self._dev(x, nx)
        if host:
            yd = torch.empty_like(xd)
            jd = torch.empty(nx, dtype=torch.float64, device=xd.device)
            ctx.invmap(xd, yd, jd)
            np.asarray(y)[:nx] = yd.cpu().numpy()
            np.asarray(jac)[:nx] = jd.cpu().numpy()
        else:
            ctx.invmap(xd, y, jac)

    def add_training_data(self, y, f, ny=-1):
        r""" Add training data ``f`` for ``y``-space points ``y`` (``_vegas.pyx:421-464``): accumulates
        ``sum_f[d, iy] += |f|`` and ``n_f[d, iy] += 1`` on the GPU. """
        if ny < 0:
            ny = y.shape[0]
        elif ny > y.shape[0]:
            raise ValueError('ny > y.shape[0]: %d > %d' % (ny, y.shape[0]))
        torch = _torch()
        ctx = self._context()
        yd, _ = self._dev(y, ny)
        fd, _ = self._dev(f, ny)
        hs = self.inc.shape[1]
        sum_f = torch.zeros((self.dim, hs), dtype=torch.float64, device=yd.device)
        n_f = torch.zeros((self.dim, hs), dtype=torch.int64, device=yd.device)
        ctx.add_training_data(yd, fd, sum_f, n_f, hs)
        self._accumulate_training(sum_f.cpu().numpy(), n_f.cpu().numpy())

    def _accumulate_training(self, sum_f, counts):
        """fold a device histogram (sums, integer counts) into sum_f / n_f (n_f starts at TINY,
        _vegas.pyx:452)"""
        if self.sum_f is None:
            shape = (self.dim, self.inc.shape[1])
            self.sum_f = np.zeros(shape, float)
            self.n_f = np.zeros(shape, float) + TINY
        self.sum_f += sum_f
        self.n_f += counts

    # ------------------------------------------------------------------ adapt
    def adapt(self, alpha=0.0, ninc=None):
        r""" Adapt grid to accumulated training data (``_vegas.pyx:467-594``); see the reference
        docstring for the meaning of ``alpha`` and ``ninc``. """
        if ninc is None:
            new_ninc = np.array(self.ninc, dtype=np.intp)
        elif np.shape(ninc) == ():
            new_ninc = np.full(self.dim, int(ninc), np.intp)
        elif len(ninc) == self.dim:
            new_ninc = np.array(ninc, np.intp)
        else:
            raise ValueError('badly formed ninc = ' + str(ninc))
        if min(new_ninc) < 1:
            raise ValueError('ninc < 1: ' + str(list(new_ninc)))
        if max(new_ninc) == 1:
            new_grid = np.empty((self.dim, 2), float)
            for d in range(self.dim):
                new_grid[d, 0] = self.grid[d, 0]
                new_grid[d, 1] = self.grid[d, self.ninc[d]]
            self.grid = new_grid
            self.inc = (new_grid[:, 1:] - new_grid[:, :1]).copy()
            self.ninc = np.array(self.dim * [1], dtype=np.intp)
            self.clear()
            self._changed()
            return
        new_grid = _lib.map_adapt(self.grid, self.ninc, self.sum_f, self.n_f, alpha, new_ninc)
        inc = np.zeros((self.dim, new_grid.shape[1] - 1), float)
        for d in range(self.dim):
            n = new_ninc[d]
            inc[d, :n] = new_grid[d, 1:n + 1] - new_grid[d, :n]
            new_grid[d, n + 1:] = new_grid[d, n]
        self.grid, self.inc, self.ninc = new_grid, inc, new_ninc
        self.clear()
        self._changed()

    def adapt_to_samples(self, x, f, nitn=5, alpha=1.0, nproc=1):
        r""" Adapt map to data ``{x, f(x)}`` (``_vegas.pyx:728-804``): repeatedly ``invmap`` the samples,
        train on ``(jac * f)**2`` and ``adapt``.  ``nproc`` is accepted for compatibility; the work
        runs on the GPU. """
        x = np.ascontiguousarray(x, dtype=float)
        if len(x.shape) != 2 or x.shape[1] != self.dim:
            raise ValueError('incompatible shape of x: {}'.format(x.shape))
        if callable(f):
            fx = np.asarray(f(x), dtype=float)
        else:
            fx = np.asarray(f, dtype=float)
        if fx.ndim != 1 or fx.shape[0] != x.shape[0]:
            raise ValueError('shape of x and f(x) mismatch: {} vs {}'.format(x.shape, fx.shape))
        # work with as many increments as the samples can resolve (pyx:783-787), then return to the
        # integrator's usual number (maxinc_axis = 1000, or the map's own if larger; pyx:803-804)
        old_ninc = max(int(max(self.ninc)), 1000)
        tmp_ninc = int(min(old_ninc, x.shape[0] / 10.))
        if tmp_ninc < 2:
            raise ValueError('not enough samples: {}'.format(x.shape[0]))
        y = np.empty(x.shape, float)
        jac = np.empty(x.shape[0], float)
        self.adapt(ninc=tmp_ninc)
        for _ in range(nitn):
            self.invmap(x, y, jac)
            self.add_training_data(y, (jac * fx) ** 2)
            self.adapt(alpha=alpha, ninc=tmp_ninc)
        if tmp_ninc != old_ninc:
            self.adapt(ninc=old_ninc)

    def show_grid(self, ngrid=40, axes=None, shrink=False, plotter=None):
        raise NotImplementedError('AdaptiveMap.show_grid (matplotlib plotting) is outside the sampling path '
                                  'this package implements; use the grid arrays directly')
