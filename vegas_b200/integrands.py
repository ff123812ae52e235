"""Built-in integrands: device functors compiled into ``libvegas_b200.so`` (``csrc/integrands.cuh``),
each with a numpy twin ``__call__(x[n, D])`` using the same constants, so the same object runs
fused on the GPU *and* as an ordinary lbatch integrand (on the reference, on the CPU oracle, or
through the unfused callback path).
"""
import collections
import ctypes
import math

import numpy as np

from . import _lib
from ._integrand import DeviceIntegrand


class Poly(DeviceIntegrand):
    """``f = c0 + sum_d c[d] * x[d]**p[d]`` (constants and polynomials for tests)."""
    fid = _lib.F_POLY

    def __init__(self, c0=0.0, c=(), p=()):
        self.c0, self.c, self.p = float(c0), [float(v) for v in c], [int(v) for v in p]

    def params(self, dim):
        q = _lib.PolyParams()
        q.c0 = self.c0
        for d in range(min(dim, len(self.c))):
            q.c[d], q.p[d] = self.c[d], self.p[d]
        return q, None

    def __call__(self, x):
        x = np.asarray(x)
        s = np.full(x.shape[0], self.c0)
        for d in range(min(x.shape[1], len(self.c))):
            t = np.ones(x.shape[0])
            for _ in range(self.p[d]):
                t = t * x[:, d]
            s = s + self.c[d] * t
        return s


class GaussMix(DeviceIntegrand):
    """``f = norm * sum_p exp(-a |x - c_p|^2)``: one Gaussian (examples/simple.py:19-23) or several
    peaks (doc/source/eg6.py:6-19)."""
    fid = _lib.F_GAUSS_MIX

    def __init__(self, centers, a=100.0, norm=1.0):
        self.centers = np.ascontiguousarray(np.atleast_2d(centers), dtype=float)
        self.a, self.norm = float(a), float(norm)

    def params(self, dim):
        if self.centers.shape[1] != dim:
            raise ValueError('GaussMix: centers have dimension %d, integrator %d' % (self.centers.shape[1], dim))
        q = _lib.GaussMixParams()
        q.npeak, q.a, q.norm = self.centers.shape[0], self.a, self.norm
        q.centers_host = self.centers.ctypes.data
        return q, self.centers

    def __call__(self, x):
        x = np.asarray(x)
        s = np.zeros(x.shape[0])
        for c in self.centers:
            dx2 = np.zeros(x.shape[0])
            for d in range(x.shape[1]):
                dx2 += (x[:, d] - c[d]) ** 2
            s += np.exp(-self.a * dx2)
        return s * self.norm


class Ridge(DeviceIntegrand):
    """``N`` Gaussians spread evenly along the diagonal from ``lo`` to ``hi`` (examples/ridge.py:18-24):
    ``f = mean_k exp(-a * sum_d (x_d - x0_k)^2) * (a/pi)^(D/2)``, ``x0 = linspace(lo, hi, N)``."""
    fid = _lib.F_RIDGE

    def __init__(self, dim, N=1000, lo=0.4, hi=0.6, a=100.0, shifted=False):
        self.dim, self.N, self.a = int(dim), int(N), float(a)
        self.shifted = bool(shifted)      # device arithmetic: axis-order sum (False) or the shifted-mean identity
        self.x0 = np.ascontiguousarray(np.linspace(lo, hi, self.N))
        self.norm = (self.a / np.pi) ** (self.dim / 2.)

    def params(self, dim):
        if dim != self.dim:
            raise ValueError('Ridge: built for %d dimensions, integrator has %d' % (self.dim, dim))
        q = _lib.RidgeParams()
        q.n, q.mode, q.a, q.norm = self.N, int(self.shifted), self.a, self.norm
        q.x0_host = self.x0.ctypes.data
        return q, self.x0

    def __call__(self, x, block=4096):
        x = np.asarray(x)
        out = np.empty(x.shape[0])
        for i in range(0, x.shape[0], block):
            xb = x[i:i + block]
            dx2 = np.zeros((xb.shape[0], self.N))
            for d in range(x.shape[1]):
                dx2 += (xb[:, d, None] - self.x0[None, :]) ** 2
            out[i:i + block] = np.average(np.exp(-self.a * dx2), axis=1) * self.norm
        return out

    def flops_per_sample(self, c_exp):
        """algorithmic fp64 flops of one evaluation (FMA = 2)"""
        if self.shifted:
            return self.N * (3 + 1 + c_exp) + 5 * self.dim + 4
        return self.N * (3 * self.dim + 2 + c_exp)


class Genz(DeviceIntegrand):
    """Genz (1984) test family on the unit hypercube.  ``kind`` in ``oscillatory``, ``product_peak``,
    ``corner_peak``, ``gaussian``, ``c0``, ``discontinuous``; ``a`` = difficulty, ``u`` = shift.
    (Not in the reference; defined here with closed-form exact values for the parity tests.)"""
    KINDS = collections.OrderedDict([
        ('oscillatory', _lib.F_GENZ_OSC), ('product_peak', _lib.F_GENZ_PRODPEAK),
        ('corner_peak', _lib.F_GENZ_CORNER), ('gaussian', _lib.F_GENZ_GAUSS), ('c0', _lib.F_GENZ_C0),
        ('discontinuous', _lib.F_GENZ_DISC)])

    def __init__(self, kind, a, u):
        self.kind = kind
        self.fid = self.KINDS[kind]
        self.a = np.asarray(a, dtype=float)
        self.u = np.asarray(u, dtype=float)

    def params(self, dim):
        if len(self.a) != dim or len(self.u) != dim:
            raise ValueError('Genz: a, u must have one entry per dimension')
        q = _lib.GenzParams()
        for d in range(dim):
            q.a[d], q.u[d] = self.a[d], self.u[d]
        return q, None

    def __call__(self, x):
        x = np.asarray(x)
        a, u = self.a, self.u
        if self.kind == 'oscillatory':
            return np.cos(2 * np.pi * u[0] + x.dot(a))
        if self.kind == 'product_peak':
            return np.prod(1.0 / (1.0 / (a * a) + (x - u) ** 2), axis=1)
        if self.kind == 'corner_peak':
            return (1.0 + x.dot(a)) ** (-(x.shape[1] + 1.0))
        if self.kind == 'gaussian':
            return np.exp(-np.sum(a * a * (x - u) ** 2, axis=1))
        if self.kind == 'c0':
            return np.exp(-np.sum(a * np.abs(x - u), axis=1))
        zero = (x[:, 0] > u[0]) | ((x[:, 1] > u[1]) if x.shape[1] > 1 else False)
        return np.where(zero, 0.0, np.exp(x.dot(a)))

    def exact(self):
        """closed-form integral over [0,1]^D"""
        a, u = self.a, self.u
        D = len(a)
        if self.kind == 'oscillatory':
            # Re[ e^{i 2 pi u0} prod_d (e^{i a_d} - 1) / (i a_d) ]
            z = np.exp(2j * np.pi * u[0])
            for ad in a:
                z *= (np.exp(1j * ad) - 1.0) / (1j * ad)
            return z.real
        if self.kind == 'product_peak':
            return float(np.prod(a * (np.arctan(a * (1 - u)) + np.arctan(a * u))))
        if self.kind == 'corner_peak':
            # inclusion-exclusion: 1/(D! prod a) * sum_{S} (-1)^{|S|} / (1 + sum_{d in S} a_d)
            tot = 0.0
            for mask in range(1 << D):
                s, bits = 1.0, 0
                for d in range(D):
                    if mask >> d & 1:
                        s += a[d]
                        bits += 1
                tot += (-1) ** bits / s
            return tot / (math.factorial(D) * float(np.prod(a)))
        if self.kind == 'gaussian':
            from scipy.special import erf
            return float(np.prod(np.sqrt(np.pi) / (2 * a) * (erf(a * (1 - u)) + erf(a * u))))
        if self.kind == 'c0':
            return float(np.prod((2 - np.exp(-a * u) - np.exp(-a * (1 - u))) / a))
        out = 1.0
        for d in range(D):
            top = min(u[d], 1.0) if d < 2 else 1.0
            out *= (np.exp(a[d] * top) - 1.0) / a[d]
        return float(out)


class PathIntegral(DeviceIntegrand):
    """Lattice path integral of a 1-d particle with periodic time (examples/path_integrand.pyx:72-142):
    ``V(x) = c2 x^2 + c4 x^4``; integration variables are ``theta_j in (-pi/2, pi/2)`` with
    ``x_j = xscale * tan(theta_j)``.  Dictionary-valued: ``'exp(-E0*T)'`` and, for every ``x0`` in
    ``x0list``, ``'exp(-E0*T) * psi(x0)**2'``."""
    fid = _lib.F_PATHINT
    shape = None

    def __init__(self, T=4.0, ndT=10, m=1.0, xscale=1.0, c2=0.5, c4=0.0, x0list=()):
        self.T, self.ndT, self.m, self.xscale = float(T), int(ndT), float(m), float(xscale)
        self.c2, self.c4 = float(c2), float(c4)
        self.x0list = np.asarray(x0list, dtype=float)
        self.nf = 1 + len(self.x0list)
        self.norm = (self.m * self.ndT / 2. / np.pi / self.T) ** (self.ndT / 2.)
        self.norm_x0 = self.norm / np.pi
        self.region = self.ndT * [[-np.pi / 2, np.pi / 2]]
        self.keys = ['exp(-E0*T)'] + (['exp(-E0*T) * psi(x0)**2'] if len(self.x0list) else [])

    def V(self, x):
        x2 = x * x
        return self.c2 * x2 + self.c4 * (x2 * x2)

    def params(self, dim):
        if dim != self.ndT:
            raise ValueError('PathIntegral: ndT=%d but integrator has %d dimensions' % (self.ndT, dim))
        if len(self.x0list) > 7:
            raise ValueError('PathIntegral: at most 7 x0 values')
        q = _lib.PathIntParams()
        q.T, q.m, q.xscale, q.c2, q.c4, q.nx0 = self.T, self.m, self.xscale, self.c2, self.c4, len(self.x0list)
        for i, v in enumerate(self.x0list):
            q.x0list[i] = v
        return q, None

    def eval_array(self, theta):
        """f[n, nf] (flat layout used by the engine)"""
        theta = np.asarray(theta)
        n, ndT = theta.shape
        a = self.T / ndT
        m_2a = self.m / 2. / a
        x = self.xscale * np.tan(theta)
        Vx = self.V(x)
        jfac = self.xscale + x ** 2 / self.xscale
        jac = self.norm * np.prod(jfac, axis=1)
        jac_x0 = self.norm_x0 * np.prod(jfac[:, 1:], axis=1)
        Smid = a * Vx[:, -1]
        for j in range(1, ndT - 1):
            Smid = Smid + (m_2a * (x[:, j + 1] - x[:, j]) ** 2 + a * Vx[:, j])
        f = np.empty((n, self.nf))
        for i in range(self.nf):
            e = x[:, 0] if i == 0 else np.full(n, self.x0list[i - 1])
            Ve = Vx[:, 0] if i == 0 else self.V(e)
            S = Smid + (m_2a * ((x[:, 1] - e) ** 2 + (e - x[:, -1]) ** 2) + a * Ve)
            f[:, i] = (jac if i == 0 else jac_x0) * np.exp(-S)
        return f

    def __call__(self, theta):
        f = self.eval_array(theta)
        ans = collections.OrderedDict()
        ans['exp(-E0*T)'] = f[:, 0]
        if self.nf > 1:
            ans['exp(-E0*T) * psi(x0)**2'] = f[:, 1:]
        return ans
