"""CPU oracle for the vegas / vegas+ hot path -- Python driver over ``vegas_oracle.c``.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``vegas_b200``.

It restates, piece by piece, what ``/root/reference/src/vegas/_vegas.pyx`` ("pyx:N") does on
the path ``Integrator.__call__`` -> ``_random_batch`` -> ``AdaptiveMap.map`` -> reduce ->
``add_training_data`` -> ``adapt``.  The scalar loops live in C (``vegas_oracle.c``, built by
``oracle/Makefile``); the integer set-up logic and the iteration driver are below.

Pinned (tests/test_oracle_vs_reference.py, tests/golden/): against the reference's own
known-answer tests and against the unmodified reference module compiled into ``oracle/_ref``
fed the same uniforms through its ``ran_array_generator`` hook.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TINY = 10.0 ** -257                                      # pyx:34
HUGE = 10.0 ** 258                                       # pyx:35
EPSILON = np.finfo(float).eps * 1e4                        # pyx:36

_lib = None


def build(force=False):
    """Compile the C restatement (gcc, ~1 s)."""
    so = os.path.join(_HERE, 'libvegas_oracle.so')
    src = os.path.join(_HERE, 'vegas_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'port'])
    return so


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int64)
        i64, i32, f64 = ctypes.c_int64, ctypes.c_int, ctypes.c_double
        L.vo_map.argtypes = [dp, ip, i64, i32, dp, dp, dp, i64]
        L.vo_jac1d.argtypes = [dp, ip, i64, i32, dp, dp, i64]
        L.vo_invmap.argtypes = [dp, ip, i64, i32, dp, dp, dp, i64]
        L.vo_add_training_data.argtypes = [ip, i32, i64, dp, dp, i64, dp, dp]
        L.vo_adapt.argtypes = [dp, ip, i64, i32, i32, dp, dp, i64, f64, ip, dp, i64]
        L.vo_alloc_neval.argtypes = [dp, i64, f64, i64, i64, ip, ip]
        L.vo_alloc_neval.restype = i64
        L.vo_stratify.argtypes = [ip, i32, i64, i64, ip, dp, dp, ip]
        L.vo_weights.argtypes = [dp, ip, i64, f64]
        L.vo_reduce_batch.argtypes = [dp, dp, i64, ip, i64, i32, i32, f64, dp, dp, dp, dp, dp, dp]
        L.vo_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.vo_philox_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint32, i32, i64, i64, ip, dp]
        for name in ('vo_map', 'vo_jac1d', 'vo_invmap', 'vo_add_training_data', 'vo_adapt',
                     'vo_stratify', 'vo_weights', 'vo_reduce_batch', 'vo_philox4x32_10',
                     'vo_philox_uniforms'):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# --------------------------------------------------------------------------- Philox
def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().vo_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def philox_uniforms(seed, itn, dim, hcube0, neval_hcube):
    """uniforms yran[n, dim] of the engine's stream for hypercubes hcube0.. (cube order)."""
    nh = _i64(neval_hcube)
    out = np.empty((int(nh.sum()), dim), float)
    lib().vo_philox_uniforms(int(seed), int(itn), dim, int(hcube0), len(nh), _ip(nh), _dp(out))
    return out


# --------------------------------------------------------------------------- AdaptiveMap
class Map(object):
    """Restatement of ``AdaptiveMap`` (pyx:39-600): grid nodes only, inc derived."""

    def __init__(self, grid, ninc=None):
        if isinstance(grid, Map):
            self.ninc = grid.ninc.copy()
            self.grid = grid.grid.copy()
        else:
            rows = [sorted(float(v) for v in g) for g in grid]          # pyx:128
            lens = [len(r) for r in rows]
            if min(lens) < 2:
                raise ValueError('grid[d] must have at least 2 elements, not %d' % min(lens))
            self.ninc = np.array(lens, dtype=np.int64) - 1
            self.grid = np.full((len(rows), max(lens)), np.nan)
            for d, r in enumerate(rows):
                self.grid[d, :len(r)] = r
        self.clear()
        if ninc is not None and not np.all(np.asarray(ninc) == self.ninc):
            if np.all(self.ninc == 1):
                self.make_uniform(ninc)                                    # pyx:134-135
            else:
                self.adapt(ninc=ninc)

    @property
    def dim(self):
        return self.grid.shape[0]

    @property
    def inc(self):
        out = np.full((self.dim, self.grid.shape[1] - 1), np.nan)
        for d in range(self.dim):
            n = self.ninc[d]
            out[d, :n] = self.grid[d, 1:n + 1] - self.grid[d, :n]
        return out

    def clear(self):
        self.sum_f = None
        self.n_f = None

    def _ninc_arg(self, ninc):
        if ninc is None:
            return self.ninc.copy()
        if np.shape(ninc) == ():
            return np.full(self.dim, int(ninc), dtype=np.int64)
        if len(ninc) != self.dim:
            raise ValueError('badly formed ninc = ' + str(ninc))
        return np.array(ninc, dtype=np.int64)

    def make_uniform(self, ninc=None):                                     # pyx:205-241
        ninc = self._ninc_arg(ninc)
        if min(ninc) < 1:
            raise ValueError('no of increments < 1 in AdaptiveMap -- %s' % str(ninc))
        g = np.full((self.dim, max(ninc) + 1), np.nan)
        for d in range(self.dim):
            g[d, :ninc[d] + 1] = np.linspace(self.grid[d, 0], self.grid[d, self.ninc[d]], ninc[d] + 1)
        self.grid, self.ninc = g, ninc
        self.clear()

    def map(self, y):
        y = _f64(y)
        x = np.empty_like(y)
        jac = np.empty(y.shape[0])
        g = _f64(self.grid)
        lib().vo_map(_dp(g), _ip(self.ninc), g.shape[1], self.dim, _dp(y), _dp(x), _dp(jac), y.shape[0])
        return x, jac

    def jac1d(self, y):
        y = _f64(y)
        out = np.empty_like(y)
        g = _f64(self.grid)
        lib().vo_jac1d(_dp(g), _ip(self.ninc), g.shape[1], self.dim, _dp(y), _dp(out), y.shape[0])
        return out

    def invmap(self, x):
        x = _f64(x)
        y = np.empty_like(x)
        jac = np.empty(x.shape[0])
        g = _f64(self.grid)
        lib().vo_invmap(_dp(g), _ip(self.ninc), g.shape[1], self.dim, _dp(x), _dp(y), _dp(jac), x.shape[0])
        return y, jac

    def add_training_data(self, y, f):
        y, f = _f64(y), _f64(f)
        if self.sum_f is None:
            shape = (self.dim, self.grid.shape[1] - 1)
            self.sum_f = np.zeros(shape)
            self.n_f = np.zeros(shape) + TINY                              # pyx:452
        lib().vo_add_training_data(_ip(self.ninc), self.dim, self.sum_f.shape[1],
                                   _dp(y), _dp(f), y.shape[0], _dp(self.sum_f), _dp(self.n_f))

    def adapt(self, alpha=0.0, ninc=None):
        new_ninc = self._ninc_arg(ninc)
        if min(new_ninc) < 1:
            raise ValueError('ninc < 1: ' + str(list(new_ninc)))
        if max(new_ninc) == 1:                                             # pyx:519-530
            g = np.empty((self.dim, 2))
            for d in range(self.dim):
                g[d] = self.grid[d, 0], self.grid[d, self.ninc[d]]
            self.grid, self.ninc = g, np.ones(self.dim, dtype=np.int64)
            self.clear()
            return
        g = _f64(self.grid)
        ng = np.empty((self.dim, max(new_ninc) + 1))
        have = self.sum_f is not None
        sf = self.sum_f if have else np.zeros((1, 1))
        nf = self.n_f if have else np.zeros((1, 1))
        lib().vo_adapt(_dp(g), _ip(self.ninc), g.shape[1], self.dim, int(have), _dp(sf), _dp(nf),
                       sf.shape[1], float(alpha), _ip(new_ninc), _dp(ng), ng.shape[1])
        self.grid, self.ninc = ng, new_ninc
        self.clear()


# --------------------------------------------------------------------------- Integrator.set
def strata(neval, dim, map_ninc, neval_frac=0.75, beta=0.75, adapt_to_errors=False,
           maxinc_axis=1000, max_mem=1e9, nstrat=None, uniform_nstrat=False, minimize_mem=False):
    """Integer set-up of ``Integrator.set`` steps 3-5 (pyx:1331-1407).

    Returns dict(neval, nstrat, ninc, nhcube, min_neval_hcube, neval_frac_eff).
    ``map_ninc`` is only used by the caller; kept for symmetry.
    """
    nf = 0 if (beta == 0 or adapt_to_errors) else neval_frac              # pyx:1331
    if nstrat is not None:
        nstrat = np.array(nstrat, dtype=np.int64)
        if len(nstrat) != dim or min(nstrat) < 1:
            raise ValueError('bad nstrat')
        nhcube = int(np.prod(nstrat))
        if neval is None:
            neval = int(2. * nhcube / (1. - nf))                           # pyx:1343
        elif neval < 2. * nhcube / (1. - nf):
            raise ValueError('neval too small')
    else:
        ns = int(abs((1 - nf) * neval / 2.) ** (1. / dim))                 # pyx:1348
        if ns < 1:
            ns = 1
        d = int((np.log((1 - nf) * neval / 2.) - dim * np.log(ns)) / np.log(1 + 1. / ns))
        if ((ns + 1) ** d * ns ** (dim - d)) > max_mem and not minimize_mem:
            raise MemoryError('work arrays larger than max_mem')
        if uniform_nstrat:
            d = 0
        nstrat = np.empty(dim, np.int64)
        nstrat[:d] = ns + 1
        nstrat[d:] = ns
    if adapt_to_errors:
        ninc = nstrat.copy()                                               # pyx:1372-1373
    else:
        ni = min(int(neval / 10.), maxinc_axis)                            # pyx:1375
        ninc = np.empty(dim, np.int64)
        for d in range(dim):
            if ni >= nstrat[d]:
                ninc[d] = int(ni / nstrat[d]) * nstrat[d]
            elif nstrat[d] <= maxinc_axis:
                ninc[d] = nstrat[d]
            else:
                nstrat[d] = int(nstrat[d] / ni) * ni
                ninc[d] = ni
    nhcube = int(np.prod(nstrat))
    if nhcube == 1:
        mnh = int(neval)
    else:
        mnh = int((1 - nf) * neval / nhcube)                               # pyx:1405
    mnh = max(mnh, 2)
    return dict(neval=int(neval), nstrat=nstrat, ninc=ninc, nhcube=nhcube,
                min_neval_hcube=mnh, neval_frac_eff=nf)


# --------------------------------------------------------------------------- the iteration
class Vegas(object):
    """Restatement of the state ``Integrator`` carries across iterations plus ONE iteration of
    ``Integrator.__call__`` (pyx:2086-2217), batch structure included (pyx:1692-1765)."""

    def __init__(self, limits, neval=1000, nstrat=None, alpha=0.5, beta=0.75, neval_frac=0.75,
                 min_neval_batch=100000, max_neval_hcube=50000, maxinc_axis=1000, max_mem=1e9,
                 adapt=True, adapt_to_errors=False, correlate_integrals=True, uniform_nstrat=False):
        self.map = limits if isinstance(limits, Map) else Map(limits)
        self.dim = self.map.dim
        self.alpha, self.beta, self.neval_frac = alpha, beta, neval_frac
        self.min_neval_batch, self.max_neval_hcube = int(min_neval_batch), int(max_neval_hcube)
        self.adapt, self.adapt_to_errors = adapt, adapt_to_errors
        self.correlate_integrals = correlate_integrals
        s = strata(neval if (nstrat is None or neval is not None) else None, self.dim, self.map.ninc,
                   neval_frac, beta, adapt_to_errors, maxinc_axis, max_mem, nstrat, uniform_nstrat)
        self.neval, self.nstrat, self.nhcube = s['neval'], s['nstrat'], s['nhcube']
        self.min_neval_hcube = s['min_neval_hcube']
        if not np.all(self.map.ninc == s['ninc']):
            self.map.adapt(ninc=s['ninc'])                                 # pyx:1385-1386
        self.sigf = np.ones(self.nhcube) if (beta >= 0 and self.nhcube > 1 and not adapt_to_errors) \
            else np.array([], float)                                       # pyx:1416-1429
        self.sum_sigf = float(self.nhcube) if len(self.sigf) else HUGE
        self.last_neval = 0
        self.neval_hcube_range = None

    def neval_sigf(self):                                                  # pyx:1657-1661
        if self.beta > 0 and self.sum_sigf > 0 and not self.adapt_to_errors:
            return self.neval_frac * self.neval / self.sum_sigf
        return 0.0

    def allocation(self):
        """neval_hcube[nhcube] for the coming iteration (pyx:1692-1706)."""
        adaptive = self.beta > 0 and self.nhcube > 1 and not self.adapt_to_errors   # pyx:1675
        rng = np.zeros(2, np.int64) + self.min_neval_hcube                 # pyx:1682
        if adaptive:
            out = np.empty(self.nhcube, np.int64)
            mx = max(self.max_neval_hcube, self.min_neval_hcube)           # pyx:1667-1669
            lib().vo_alloc_neval(_dp(self.sigf), self.nhcube, self.neval_sigf(),
                                 self.min_neval_hcube, mx, _ip(out), _ip(rng))
        else:
            out = np.full(self.nhcube, int(self.neval / self.nhcube), np.int64)     # pyx:1662,1705
        return out, rng

    def batches(self, neval_hcube):
        """(hcube0, hcube1) ranges exactly as the reference cuts them (pyx:1708-1710)."""
        out, acc, base = [], 0, 0
        csum = np.cumsum(neval_hcube)
        h = 0
        while h < self.nhcube:
            # first h' >= h with csum[h'] - csum[base-1] >= min_neval_batch
            start = csum[base - 1] if base > 0 else 0
            hp = int(np.searchsorted(csum, start + self.min_neval_batch, side='left'))
            hp = min(hp, self.nhcube - 1)
            out.append((base, hp + 1))
            base = hp + 1
            h = base
        return out

    def iterate(self, fcn, uniforms, nf=None):
        """One iteration.  ``fcn(x[n,D]) -> f[n] or f[n,nf]``; ``uniforms(hcube0, neval_hcube_batch)
        -> yran[n,D]``.  Returns (mean[nf], var) and updates sigf / sum_sigf / training data.
        The map is NOT adapted here (call ``adapt_map()``), so tests can inspect sum_f / n_f."""
        L = lib()
        neval_hcube, rng = self.allocation()
        self.neval_hcube_range = rng
        self.last_neval = int(neval_hcube.sum())
        dv_y = 1. / self.nhcube                                            # pyx:1651
        update_sigf = (self.beta > 0 and self.nhcube > 1 and self.adapt and not self.adapt_to_errors)
        mean = var = None
        sum_sigf = np.zeros(1)
        self.samples = []
        for h0, h1 in self.batches(neval_hcube):
            nh = np.ascontiguousarray(neval_hcube[h0:h1])
            n = int(nh.sum())
            yran = _f64(uniforms(h0, nh))
            assert yran.shape == (n, self.dim)
            y = np.empty((n, self.dim))
            hc = np.empty(n, np.int64)
            L.vo_stratify(_ip(self.nstrat), self.dim, h0, len(nh), _ip(nh), _dp(yran), _dp(y), _ip(hc))
            x, jac = self.map.map(y)
            L.vo_weights(_dp(jac), _ip(nh), len(nh), dv_y)
            fx = np.asarray(fcn(x), dtype=float)
            fx = _f64(fx.reshape(n, -1))
            if np.any(np.isnan(fx)):
                raise ValueError('integrand evaluates to nan')             # pyx:2133-2134
            k = fx.shape[1]
            if mean is None:
                mean = np.zeros(k)
                var = np.zeros((k, k)) if self.correlate_integrals else np.zeros(k)
            fdv2 = np.empty(n)
            sigf_view = np.ascontiguousarray(self.sigf[h0:h1]) if update_sigf else np.zeros(1)
            sigf2 = np.empty(len(nh))
            L.vo_reduce_batch(_dp(jac), _dp(fx), k, _ip(nh), len(nh), int(self.correlate_integrals),
                              int(update_sigf), float(self.beta), _dp(mean), _dp(var),
                              _dp(sigf_view), _dp(sum_sigf), _dp(fdv2), _dp(sigf2))
            if update_sigf:
                self.sigf[h0:h1] = sigf_view
            if self.adapt_to_errors and self.adapt:                        # pyx:2187-2193
                last = np.cumsum(nh) - 1
                self.map.add_training_data(y[last], sigf2)
            elif self.adapt and self.alpha > 0:                            # pyx:2196-2197
                self.map.add_training_data(y, fdv2)
            self.samples.append((x, y, jac, hc, fx))
        if self.correlate_integrals:
            var = np.tril(var) + np.tril(var, -1).T                        # pyx:2199-2202
        if self.beta > 0 and not self.adapt_to_errors and self.adapt:      # pyx:2209-2215
            if sum_sigf[0] > 0:
                self.sum_sigf = float(sum_sigf[0])
            else:
                self.sigf[:] = 1.
                self.sum_sigf = float(len(self.sigf))
        return mean, var

    def adapt_map(self):
        if self.alpha > 0 and self.adapt:                                  # pyx:2216-2217
            self.map.adapt(alpha=self.alpha)


# --------------------------------------------------------------------------- result averaging
def wavg(means, variances):
    """Weighted running average of scalar estimates (pyx:2392-2403) + chi2/dof/Q (pyx:2344-2377)."""
    from scipy.special import gammaincc
    w = [1. / (v if v > TINY else TINY) for v in variances]
    var = 1. / np.sum(w)
    mean = np.sum([wi * mi for wi, mi in zip(w, means)]) * var
    chi2 = 0.0
    if len(means) > 1:
        for m, wi in zip(means, w):
            chi2 += (mean - m) ** 2 * wi
    dof = len(means) - 1
    Q = float(gammaincc(dof / 2., chi2 / 2.)) if dof > 0 and chi2 >= 0 else float('nan')
    return mean, math.sqrt(var), chi2, dof, Q


def uavg(means, variances):
    """Unweighted average (adapt=False; pyx:2405-2410, 2354-2356)."""
    from scipy.special import gammaincc
    n = len(means)
    mean = np.sum(means) / n
    var = np.sum(variances) / n ** 2
    chi2 = float(np.sum([(m - mean) ** 2 for m in means]) / (np.sum(variances) / n)) if n > 1 else 0.0
    dof = n - 1
    Q = float(gammaincc(dof / 2., chi2 / 2.)) if dof > 0 and chi2 >= 0 else float('nan')
    return mean, math.sqrt(var), chi2, dof, Q


# --------------------------------------------------------------------------- restratify
def profile_integrand(vmap, f, ndy):
    """The auxiliary integrand of ``vegas.restratify`` (src/vegas/__init__.py:1390-1419):
    component 0 is ``I = f(x)[:, 0]``; component ``1 + mu*ndy + i`` is ``I`` where
    ``yst[i] <= y[mu] <= yst[i+1]`` (closed on both sides) with ``y = map.invmap(x)``, else 0."""
    yst = np.linspace(0, 1, ndy + 1)                                       # __init__.py:1321

    def fcn(x):
        x = np.asarray(x, dtype=float)
        n, dim = x.shape
        y, _ = vmap.invmap(x)                                              # __init__.py:1402
        I = np.asarray(f(x), dtype=float).reshape(n, -1)[:, 0]             # __init__.py:1406
        out = np.zeros((n, 1 + dim * ndy))
        out[:, 0] = I
        for mu in range(dim):
            for i in range(ndy):
                idx = (yst[i] <= y[:, mu]) & (y[:, mu] <= yst[i + 1])      # __init__.py:1411
                out[idx, 1 + mu * ndy + i] = I[idx]
        return out
    return fcn


def restratify_weights(I, dI, ndy):
    """weight[mu] = sum_i (dI[mu][i] - I/ndy)^2 * ndy   (__init__.py:1340-1344), on mean values"""
    dI = np.asarray(dI, dtype=float)
    return np.sum((dI - I / ndy) ** 2, axis=1) * ndy


def restratify_nstrat(old_nstrat, weight, gamma=1.0, below_avg_nstrat=None):
    """new strata per axis (__init__.py:1349-1365): proportional to the weights at constant
    geometric-mean strata, assigned smallest first; what rounding (or ``below_avg_nstrat``) takes
    from an axis is handed to the axes not yet assigned."""
    old = np.asarray(old_nstrat)
    dim = len(old)
    weight = np.asarray(weight, dtype=float)
    w_avg = np.average(weight)
    w_gm = np.prod(weight) ** (1 / dim)
    nstrat_gm = np.prod(old) ** (1 / dim)
    nstrat = old * ((weight / w_gm) * nstrat_gm / old) ** gamma
    nleft = dim
    musort = np.array(np.argsort(nstrat))
    new = np.array(old)
    for mu in musort:
        new[mu] = nstrat[mu] if nstrat[mu] > 1 else 1
        if below_avg_nstrat and weight[mu] < w_avg:
            new[mu] = below_avg_nstrat
        nleft -= 1
        if nleft > 0:
            nstrat[musort[-nleft:]] *= (nstrat[mu] / new[mu]) ** (1 / nleft)
            nstrat[mu] = new[mu]
    return new
